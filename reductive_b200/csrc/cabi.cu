// cabi.cu — implementation of include/reductive_b200.h: handle management, layout normalisation at the
// boundary, host<->device pipelines and the k-means training loop.  All arithmetic happens in the CUDA
// kernels of this directory; there is no CPU compute path here (only copies, packing of odd strides and
// the best-of-attempts selection on M floats).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "encode_tc.cuh"
#include "project_tc.cuh"

namespace rb {

std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_encode_algo{RB_ENCODE_AUTO};
static std::atomic<int> g_project_algo{RB_PROJECT_AUTO};
static std::atomic<int> g_kmeans_ordered{1};
static thread_local char t_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

bool kmeans_stream_enabled() { return g_kmeans_ordered.load() == 3; }
int kmeans_update_mode() { return g_kmeans_ordered.load(); }

bool debug_sync_enabled()
{
    static const bool on = getenv("RB_DEBUG_SYNC") != nullptr;
    return on;
}

void debug_sync_report(const char *file, int line)
{
    fprintf(stderr, "[rb sync] launched at %s:%d ...", file, line);
    fflush(stderr);
    const cudaError_t e = cudaDeviceSynchronize();
    fprintf(stderr, " %s\n", e == cudaSuccess ? "done" : cudaGetErrorString(e));
    fflush(stderr);
}

static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {};

cudaError_t pool_malloc(void **p, size_t bytes, cudaStream_t stream)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMallocAsync(p, bytes, stream);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (!g_pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            e = cudaMemPoolCreate(&g_pools[dev], &props);
            if (e != cudaSuccess) return e;
            unsigned long long keep = ~0ull;  // keep everything until rb_release_scratch
            cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool = g_pools[dev];
    }
    return cudaMallocFromPoolAsync(p, bytes, pool, stream);
}

static rb_status fail(rb_status s, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return s;
}

static rb_status require_device()
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return fail(RB_ERR_NO_DEVICE,
                    "no CUDA device visible (%s); reductive_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    static std::once_flag once[64];
    int dev = 0;
    RB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64) {
        std::call_once(once[dev], [dev]() {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t thr = UINT64_MAX;  // keep freed workspaces cached in the pool
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        });
    }
    return RB_OK;
}

// Stream-ordered workspace that frees itself.
struct Workspace {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    rb_status alloc(size_t bytes, cudaStream_t stream)
    {
        s = stream;
        RB_CUDA_TRY(pool_malloc((void **)&p, bytes ? bytes : 1, stream));
        return RB_OK;
    }
    ~Workspace()
    {
        if (p) cudaFreeAsync(p, s);
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace rb

using namespace rb;

struct rb_pq {
    int device = 0;
    size_t M = 0, k = 0, dsub = 0, d = 0;
    float *q_dev = nullptr;     // [M,k,dsub]
    float *cs_dev = nullptr;    // [M,k]
    float *proj_dev = nullptr;  // [d,d] or null
    float *projt_dev = nullptr; // its transpose: decode multiplies by R^T as a plain row-major matrix (same FMA order)
    TensorOperands tc;          // bf16-split codebook for the tcgen05 path (may be empty)
    ProjTensorOperands ptc_enc, ptc_dec;  // R / R^T limbs for the tcgen05 rotation (may be empty)
    float q_absmax = 0.f;        // max |centroid component| (bounds every gathered value); < 0: non-finite codebook
    std::vector<float> q_host, proj_host;
    // rb_pq_create_multi: replicas of this quantizer on the other devices (owned by this handle); host-memory batch
    // calls split their rows over this device and the replicas' devices
    std::vector<rb_pq *> peers;
    DeviceCodebook cb() const { return DeviceCodebook{q_dev, cs_dev, M, k, dsub}; }
};

static bool valid_width(int w) { return w == 1 || w == 2 || w == 4 || w == 8; }

// ---------------------------------------------------------------------------------------------------
// encode on device-resident, row-major-with-pitch data
// ---------------------------------------------------------------------------------------------------
static rb_status encode_device(const DeviceCodebook &cb, const TensorOperands *tc, const float *x, size_t n,
                               ptrdiff_t ldx, int seq_norm, void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs,
                               cudaStream_t stream)
{
    int algo = g_encode_algo.load();
    const bool tc_ok = tc != nullptr && tc->ready() && !seq_norm && tensor_call_supported(cb, x, n, ldx);
    if (algo == RB_ENCODE_TENSOR && !tc_ok)
        return fail(RB_ERR_UNSUPPORTED, "tensor encode path does not cover this shape (k=%zu, dsub=%zu)", cb.k, cb.dsub);
    if (algo == RB_ENCODE_AUTO) algo = tc_ok ? RB_ENCODE_TENSOR : RB_ENCODE_EXACT;
    if (algo == RB_ENCODE_TENSOR)
        return launch_encode_tensor(cb, *tc, x, n, ldx, codes, code_width, crs, ccs, stream);
    return launch_encode_exact(cb, x, n, ldx, codes, code_width, crs, ccs, seq_norm, stream);
}

// rows per workspace chunk so that a [rows, d] f32 temporary stays around 2 GiB
static size_t workspace_rows(size_t n, size_t d)
{
    size_t rows = ((size_t)2 << 30) / (d * sizeof(float));
    if (rows < 1024) rows = 1024;
    return rows < n ? rows : n;
}

static rb_status quantize_batch_device(const rb_pq *pq, const float *x, size_t n, ptrdiff_t rs, ptrdiff_t cs,
                                       void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs, cudaStream_t stream)
{
    const size_t d = pq->d;
    const DeviceCodebook cb = pq->cb();
    if (!pq->proj_dev && cs == 1) {  // primitives.rs:96: column slices of a standard-layout view
        return encode_device(cb, &pq->tc, x, n, rs, 0, codes, code_width, crs, ccs, stream);
    }
    const size_t chunk = workspace_rows(n, d);
    Workspace ws;
    RB_TRY(ws.alloc(chunk * d * sizeof(float), stream));
    for (size_t r0 = 0; r0 < n; r0 += chunk) {
        const size_t rows = n - r0 < chunk ? n - r0 : chunk;
        const float *xs = x + (ptrdiff_t)r0 * rs;
        int seq_norm = 0;
        if (pq->proj_dev) {  // pq.rs:276: rx = x.dot(projection), a fresh standard-layout array
            // Large aligned batches: tensor-core rotation, then the tensor encode decides what is not close and
            // re-rotates exactly what is (codes stay bit-exact, encode_tc.cuh RotatedInput).
            const int palgo = g_project_algo.load();
            const bool ptc_ok = cs == 1 && g_encode_algo.load() != RB_ENCODE_EXACT && pq->tc.ready() &&
                                tensor_call_supported(cb, ws.as<float>(), rows, (ptrdiff_t)d) &&
                                rotated_recheck_supported(cb, d) && project_tensor_rowerr_supported(pq->ptc_enc) &&
                                project_tensor_call_supported(pq->ptc_enc, xs, rows, rs, ws.as<float>(), (ptrdiff_t)d);
            if (palgo == RB_PROJECT_TENSOR && !ptc_ok)
                return fail(RB_ERR_UNSUPPORTED, "tensor projection does not cover this call (d=%zu, n=%zu)", d, rows);
            if (palgo != RB_PROJECT_EXACT && ptc_ok) {
                Workspace aux;  // [rows] per-row error bound of the rotation | scale, |x|max bits, counter, pad
                RB_TRY(aux.alloc((rows + 4) * sizeof(float), stream));
                float *rowerr = aux.as<float>(), *scale4 = aux.as<float>() + rows;
                RB_TRY(launch_project_sample_scale(xs, rows, d, rs, scale4, stream));
                RB_TRY(launch_project_tensor(pq->ptc_enc, xs, rows, rs, scale4, 0.f, ws.as<float>(), (ptrdiff_t)d, rowerr,
                                             stream));
                const RotatedInput rot{xs, rs, pq->proj_dev, d, rowerr, scale4, project_tensor_error_floor(pq->ptc_enc)};
                char *cdst = reinterpret_cast<char *>(codes) + (ptrdiff_t)r0 * crs * code_width;
                RB_TRY(launch_encode_tensor(cb, pq->tc, ws.as<float>(), rows, (ptrdiff_t)d, cdst, code_width, crs, ccs,
                                            stream, &rot));
                continue;
            }
            RB_TRY(launch_project(xs, rows, d, rs, cs, pq->proj_dev, 0, ws.as<float>(), stream));
        } else {  // non-unit column stride: pack; the reference's row norms take the sequential dot then
            RB_TRY(launch_pack_rows(xs, rows, d, rs, cs, ws.as<float>(), stream));
            seq_norm = (d > 1 && pq->dsub > 1) ? 1 : 0;
        }
        char *cdst = reinterpret_cast<char *>(codes) + (ptrdiff_t)r0 * crs * code_width;
        RB_TRY(encode_device(cb, &pq->tc, ws.as<float>(), rows, (ptrdiff_t)d, seq_norm, cdst, code_width, crs, ccs,
                             stream));
    }
    return RB_OK;
}

static rb_status reconstruct_batch_device(const rb_pq *pq, const void *codes, int code_width, size_t n, ptrdiff_t crs,
                                          ptrdiff_t ccs, float *out, ptrdiff_t ors, ptrdiff_t ocs, int *err_flag,
                                          cudaStream_t stream)
{
    const size_t d = pq->d;
    const DeviceCodebook cb = pq->cb();
    if (!pq->proj_dev && ocs == 1)
        return launch_gather(cb, codes, code_width, n, crs, ccs, out, ors, err_flag, stream);
    const size_t chunk = workspace_rows(n, d);
    Workspace y, z;
    RB_TRY(y.alloc(chunk * d * sizeof(float), stream));
    const bool direct = pq->proj_dev && ocs == 1 && ors == (ptrdiff_t)d;
    if (pq->proj_dev && !direct) RB_TRY(z.alloc(chunk * d * sizeof(float), stream));
    for (size_t r0 = 0; r0 < n; r0 += chunk) {
        const size_t rows = n - r0 < chunk ? n - r0 : chunk;
        const char *csrc = reinterpret_cast<const char *>(codes) + (ptrdiff_t)r0 * crs * code_width;
        float *odst = out + (ptrdiff_t)r0 * ors;
        RB_TRY(launch_gather(cb, csrc, code_width, rows, crs, ccs, y.as<float>(), (ptrdiff_t)d, err_flag, stream));
        if (pq->proj_dev) {  // pq.rs:323-326: reconstructions.dot(&projection.t()), then assign
            float *pdst = direct ? odst : z.as<float>();
            const ptrdiff_t pld = direct ? ors : (ptrdiff_t)d;
            // the rotated reconstruction is allowed 1e-5 (north_star): large batches take the tensor-core GEMM
            const int palgo = g_project_algo.load();
            const bool ptc_ok = pq->q_absmax >= 0.f &&
                                project_tensor_call_supported(pq->ptc_dec, y.as<float>(), rows, (ptrdiff_t)d, pdst, pld);
            if (palgo == RB_PROJECT_TENSOR && !ptc_ok)
                return fail(RB_ERR_UNSUPPORTED, "tensor projection does not cover this call (d=%zu, n=%zu)", d, rows);
            if (palgo != RB_PROJECT_EXACT && ptc_ok) {
                RB_TRY(launch_project_tensor(pq->ptc_dec, y.as<float>(), rows, (ptrdiff_t)d, nullptr,
                                             project_scale_for_absmax(pq->q_absmax), pdst, pld, nullptr, stream));
            } else {
                RB_TRY(launch_project(y.as<float>(), rows, d, (ptrdiff_t)d, 1, pq->projt_dev, 0, pdst, stream));
            }
            if (!direct) RB_TRY(launch_unpack_rows(z.as<float>(), rows, d, odst, ors, ocs, stream));
        } else {
            RB_TRY(launch_unpack_rows(y.as<float>(), rows, d, odst, ors, ocs, stream));
        }
    }
    return RB_OK;
}

// ---------------------------------------------------------------------------------------------------
// host-memory pipelines: chunked, double-buffered H2D -> kernels -> D2H
// ---------------------------------------------------------------------------------------------------
namespace {

struct HostPipe {
    cudaStream_t st[2] = {nullptr, nullptr};
    rb_status init()
    {
        for (int i = 0; i < 2; i++) RB_CUDA_TRY(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        return RB_OK;
    }
    ~HostPipe()
    {
        for (int i = 0; i < 2; i++)
            if (st[i]) {
                cudaStreamSynchronize(st[i]);
                cudaStreamDestroy(st[i]);
            }
    }
};

size_t host_chunk_rows(size_t n, size_t row_bytes)
{
    size_t rows = ((size_t)64 << 20) / (row_bytes ? row_bytes : 1);
    if (rows < 4096) rows = 4096;
    return rows < n ? rows : n;
}

// dense copy of a strided host matrix of `elem` byte elements into dst [rows, cols]
void host_pack(const void *src, size_t rows, size_t cols, ptrdiff_t rs, ptrdiff_t cs, size_t elem, void *dst)
{
    const char *s = reinterpret_cast<const char *>(src);
    char *d = reinterpret_cast<char *>(dst);
    for (size_t r = 0; r < rows; r++)
        for (size_t c = 0; c < cols; c++)
            memcpy(d + (r * cols + c) * elem, s + ((ptrdiff_t)r * rs + (ptrdiff_t)c * cs) * (ptrdiff_t)elem, elem);
}

void host_unpack(const void *src, size_t rows, size_t cols, void *dst, ptrdiff_t rs, ptrdiff_t cs, size_t elem)
{
    const char *s = reinterpret_cast<const char *>(src);
    char *d = reinterpret_cast<char *>(dst);
    for (size_t r = 0; r < rows; r++)
        for (size_t c = 0; c < cols; c++)
            memcpy(d + ((ptrdiff_t)r * rs + (ptrdiff_t)c * cs) * (ptrdiff_t)elem, s + (r * cols + c) * elem, elem);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// pageable host memory: library-owned pinned staging
// ---------------------------------------------------------------------------------------------------
namespace {

// An ndarray caller hands ordinary (pageable) memory; a DMA from it is staged by the driver through a small pinned
// buffer at a fraction of the PCIe rate and blocks the calling thread.  The host pipelines therefore copy pageable
// chunks through pinned buffers of their own (several host threads per copy), which keeps the H2D / D2H copies
// asynchronous and at the pinned rate.  Buffers are cached per host thread and device and live until the thread ends.
bool host_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct PinnedSet {
    struct Buf {
        void *p = nullptr;
        size_t bytes = 0;
    };
    Buf bufs[4];  // in[0], in[1], out[0], out[1]
    void *get(int i, size_t bytes)
    {
        Buf &b = bufs[i];
        if (b.bytes < bytes) {
            if (b.p) cudaFreeHost(b.p);
            b.p = nullptr;
            b.bytes = 0;
            if (cudaHostAlloc(&b.p, bytes, cudaHostAllocPortable) != cudaSuccess) {
                (void)cudaGetLastError();
                b.p = nullptr;
                return nullptr;
            }
            b.bytes = bytes;
        }
        return b.p;
    }
};

// Process-wide pool of staging sets (pinned allocations cost milliseconds: they are kept for the life of the
// process and handed from call to call; concurrent calls each take their own set).
struct PinnedLease {
    PinnedSet *set = nullptr;
    PinnedLease()
    {
        std::lock_guard<std::mutex> lock(mu());
        if (!idle().empty()) {
            set = idle().back();
            idle().pop_back();
        } else {
            set = new PinnedSet();
        }
    }
    ~PinnedLease()
    {
        std::lock_guard<std::mutex> lock(mu());
        idle().push_back(set);
    }
    void *get(int i, size_t bytes) { return set->get(i, bytes); }
    static std::mutex &mu()
    {
        static std::mutex m;
        return m;
    }
    static std::vector<PinnedSet *> &idle()
    {
        static std::vector<PinnedSet *> v;
        return v;
    }
};

std::atomic<int> g_copy_threads{4};

// dst[rows, row_bytes] (dense) <-> a host matrix with row pitch `pitch` bytes, split over a few host threads
void par_copy_rows(char *dst, size_t dpitch, const char *src, size_t spitch, size_t row_bytes, size_t rows)
{
    const size_t total = rows * row_bytes;
    int nt = g_copy_threads.load();
    if (total < ((size_t)4 << 20) || nt <= 1) nt = 1;
    auto work = [=](size_t r0, size_t r1) {
        if (dpitch == row_bytes && spitch == row_bytes) {
            memcpy(dst + r0 * row_bytes, src + r0 * row_bytes, (r1 - r0) * row_bytes);
        } else {
            for (size_t r = r0; r < r1; r++) memcpy(dst + r * dpitch, src + r * spitch, row_bytes);
        }
    };
    if (nt == 1) {
        work(0, rows);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (rows + (size_t)nt - 1) / (size_t)nt;
    for (int t = 1; t < nt; t++) {
        const size_t r0 = per * (size_t)t < rows ? per * (size_t)t : rows, r1 = r0 + per < rows ? r0 + per : rows;
        if (r0 < r1) th.emplace_back(work, r0, r1);
    }
    work(0, per < rows ? per : rows);
    for (auto &t : th) t.join();
}

// Run `body(replica, first row, rows)` for contiguous row blocks over the quantizer's devices, one host thread each.
rb_status over_devices(const rb_pq *pq, size_t n, const std::function<rb_status(const rb_pq *, size_t, size_t)> &body)
{
    const size_t D = 1 + pq->peers.size();
    if (D == 1 || n < 2 * D) {
        RB_CUDA_TRY(cudaSetDevice(pq->device));
        return body(pq, 0, n);
    }
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<rb_status> st(D, RB_OK);
    std::vector<std::string> err(D);
    std::vector<std::thread> th;
    const size_t per = ((n + D - 1) / D + 127) / 128 * 128;  // whole 128-row tiles per device
    for (size_t i = 0; i < D; i++) {
        const size_t r0 = per * i < n ? per * i : n, r1 = r0 + per < n ? r0 + per : n;
        if (r0 >= r1) continue;
        const rb_pq *rep = i == 0 ? pq : pq->peers[i - 1];
        th.emplace_back([&, i, rep, r0, r1]() {
            if (cudaSetDevice(rep->device) != cudaSuccess) {
                st[i] = RB_ERR_CUDA;
                err[i] = "cudaSetDevice failed";
                return;
            }
            st[i] = body(rep, r0, r1 - r0);
            if (st[i] != RB_OK) err[i] = rb_last_error_message();
        });
    }
    for (auto &t : th) t.join();
    cudaSetDevice(prev);
    for (size_t i = 0; i < D; i++)
        if (st[i] != RB_OK) {
            set_error("device %d: %s", i == 0 ? pq->device : pq->peers[i - 1]->device, err[i].c_str());
            return st[i];
        }
    return RB_OK;
}

}  // namespace

static rb_status quantize_batch_host(const rb_pq *pq, const float *x, size_t n, ptrdiff_t rs, ptrdiff_t cs,
                                     void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs)
{
    const size_t d = pq->d, M = pq->M, cw = (size_t)code_width;
    const bool in2d = (cs == 1 && rs >= (ptrdiff_t)d);
    const bool out2d = (ccs == 1 && crs >= (ptrdiff_t)M);
    // DMA straight from / to the caller's memory only when it is pinned; otherwise through pinned staging buffers
    const bool in_direct = in2d && host_is_pinned(x), out_direct = out2d && host_is_pinned(codes);
    const size_t chunk = host_chunk_rows(n, d * sizeof(float));
    HostPipe pipe;
    RB_TRY(pipe.init());
    Workspace xin[2], cout[2];
    PinnedLease t_pinned;
    float *sin[2] = {nullptr, nullptr};
    unsigned char *sout[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        RB_TRY(xin[i].alloc(chunk * d * sizeof(float), pipe.st[i]));
        RB_TRY(cout[i].alloc(chunk * M * cw, pipe.st[i]));
        if (!in_direct) sin[i] = reinterpret_cast<float *>(t_pinned.get(i, chunk * d * sizeof(float)));
        if (!out_direct) sout[i] = reinterpret_cast<unsigned char *>(t_pinned.get(2 + i, chunk * M * cw));
        if ((!in_direct && !sin[i]) || (!out_direct && !sout[i])) return fail(RB_ERR_CUDA, "cudaHostAlloc of a staging buffer failed");
    }
    // a strided host view keeps its meaning: without a projection the reference's row norms take the
    // sequential dot for non-unit column strides (see oracle.c); packing does not change that.
    const bool seq_norm = !pq->proj_dev && cs != 1 && d > 1 && pq->dsub > 1;
    size_t pending_r0[2] = {0, 0}, pending_rows[2] = {0, 0};
    auto drain = [&](int slot) -> rb_status {  // the slot's previous chunk is complete; deliver staged codes
        RB_CUDA_TRY(cudaStreamSynchronize(pipe.st[slot]));
        if (pending_rows[slot]) {
            char *dst = reinterpret_cast<char *>(codes) + (ptrdiff_t)pending_r0[slot] * crs * (ptrdiff_t)cw;
            if (out2d)
                par_copy_rows(dst, (size_t)crs * cw, reinterpret_cast<const char *>(sout[slot]), M * cw, M * cw, pending_rows[slot]);
            else
                host_unpack(sout[slot], pending_rows[slot], M, dst, crs, ccs, cw);
        }
        pending_rows[slot] = 0;
        return RB_OK;
    };
    int slot = 0;
    for (size_t r0 = 0; r0 < n; r0 += chunk, slot ^= 1) {
        const size_t rows = n - r0 < chunk ? n - r0 : chunk;
        cudaStream_t st = pipe.st[slot];
        RB_TRY(drain(slot));
        float *xd = xin[slot].as<float>();
        const float *xsrc = x + (ptrdiff_t)r0 * rs;
        if (in_direct && rs == (ptrdiff_t)d) {  // contiguous: one linear DMA (2-D copies of 1.2 KB rows run at ~1/5 speed)
            RB_CUDA_TRY(cudaMemcpyAsync(xd, xsrc, rows * d * sizeof(float), cudaMemcpyHostToDevice, st));
        } else if (in_direct) {
            RB_CUDA_TRY(cudaMemcpy2DAsync(xd, d * sizeof(float), xsrc, (size_t)rs * sizeof(float), d * sizeof(float), rows,
                                          cudaMemcpyHostToDevice, st));
        } else {
            if (in2d)
                par_copy_rows(reinterpret_cast<char *>(sin[slot]), d * sizeof(float), reinterpret_cast<const char *>(xsrc),
                              (size_t)rs * sizeof(float), d * sizeof(float), rows);
            else
                host_pack(xsrc, rows, d, rs, cs, sizeof(float), sin[slot]);
            RB_CUDA_TRY(cudaMemcpyAsync(xd, sin[slot], rows * d * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        if (seq_norm)
            RB_TRY(encode_device(pq->cb(), &pq->tc, xd, rows, (ptrdiff_t)d, 1, cout[slot].p, code_width, (ptrdiff_t)M, 1, st));
        else  // the same path as device-resident batches (tensor rotation + tensor encode where the shape allows)
            RB_TRY(quantize_batch_device(pq, xd, rows, (ptrdiff_t)d, 1, cout[slot].p, code_width, (ptrdiff_t)M, 1, st));
        char *cdst = reinterpret_cast<char *>(codes) + (ptrdiff_t)r0 * crs * (ptrdiff_t)cw;
        if (out_direct && crs == (ptrdiff_t)M) {
            RB_CUDA_TRY(cudaMemcpyAsync(cdst, cout[slot].p, rows * M * cw, cudaMemcpyDeviceToHost, st));
        } else if (out_direct) {
            RB_CUDA_TRY(cudaMemcpy2DAsync(cdst, (size_t)crs * cw, cout[slot].p, M * cw, M * cw, rows, cudaMemcpyDeviceToHost, st));
        } else {
            RB_CUDA_TRY(cudaMemcpyAsync(sout[slot], cout[slot].p, rows * M * cw, cudaMemcpyDeviceToHost, st));
            pending_r0[slot] = r0;
            pending_rows[slot] = rows;
        }
    }
    for (int i = 0; i < 2; i++) RB_TRY(drain(i));
    return RB_OK;
}

static rb_status reconstruct_batch_host(const rb_pq *pq, const void *codes, int code_width, size_t n, ptrdiff_t crs,
                                        ptrdiff_t ccs, float *out, ptrdiff_t ors, ptrdiff_t ocs)
{
    const size_t d = pq->d, M = pq->M, cw = (size_t)code_width;
    const bool in2d = (ccs == 1 && crs >= (ptrdiff_t)M);
    const bool out2d = (ocs == 1 && ors >= (ptrdiff_t)d);
    const bool in_direct = in2d && host_is_pinned(codes), out_direct = out2d && host_is_pinned(out);
    const size_t chunk = host_chunk_rows(n, d * sizeof(float));
    HostPipe pipe;
    RB_TRY(pipe.init());
    Workspace cin[2], yout[2], flag;
    PinnedLease t_pinned;
    unsigned char *sin[2] = {nullptr, nullptr};
    float *sout[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        RB_TRY(cin[i].alloc(chunk * M * cw, pipe.st[i]));
        RB_TRY(yout[i].alloc(chunk * d * sizeof(float), pipe.st[i]));
        if (!in_direct) sin[i] = reinterpret_cast<unsigned char *>(t_pinned.get(i, chunk * M * cw));
        if (!out_direct) sout[i] = reinterpret_cast<float *>(t_pinned.get(2 + i, chunk * d * sizeof(float)));
        if ((!in_direct && !sin[i]) || (!out_direct && !sout[i])) return fail(RB_ERR_CUDA, "cudaHostAlloc of a staging buffer failed");
    }
    RB_TRY(flag.alloc(sizeof(int), pipe.st[0]));
    RB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int), pipe.st[0]));
    RB_CUDA_TRY(cudaStreamSynchronize(pipe.st[0]));
    size_t pending_r0[2] = {0, 0}, pending_rows[2] = {0, 0};
    auto drain = [&](int slot) -> rb_status {
        RB_CUDA_TRY(cudaStreamSynchronize(pipe.st[slot]));
        if (pending_rows[slot]) {
            float *dst = out + (ptrdiff_t)pending_r0[slot] * ors;
            if (out2d)
                par_copy_rows(reinterpret_cast<char *>(dst), (size_t)ors * sizeof(float), reinterpret_cast<const char *>(sout[slot]),
                              d * sizeof(float), d * sizeof(float), pending_rows[slot]);
            else
                host_unpack(sout[slot], pending_rows[slot], d, dst, ors, ocs, sizeof(float));
        }
        pending_rows[slot] = 0;
        return RB_OK;
    };
    int slot = 0;
    for (size_t r0 = 0; r0 < n; r0 += chunk, slot ^= 1) {
        const size_t rows = n - r0 < chunk ? n - r0 : chunk;
        cudaStream_t st = pipe.st[slot];
        RB_TRY(drain(slot));
        const char *csrc = reinterpret_cast<const char *>(codes) + (ptrdiff_t)r0 * crs * (ptrdiff_t)cw;
        if (in_direct && crs == (ptrdiff_t)M) {
            RB_CUDA_TRY(cudaMemcpyAsync(cin[slot].p, csrc, rows * M * cw, cudaMemcpyHostToDevice, st));
        } else if (in_direct) {
            RB_CUDA_TRY(cudaMemcpy2DAsync(cin[slot].p, M * cw, csrc, (size_t)crs * cw, M * cw, rows, cudaMemcpyHostToDevice, st));
        } else {
            if (in2d)
                par_copy_rows(reinterpret_cast<char *>(sin[slot]), M * cw, csrc, (size_t)crs * cw, M * cw, rows);
            else
                host_pack(csrc, rows, M, crs, ccs, cw, sin[slot]);
            RB_CUDA_TRY(cudaMemcpyAsync(cin[slot].p, sin[slot], rows * M * cw, cudaMemcpyHostToDevice, st));
        }
        RB_TRY(reconstruct_batch_device(pq, cin[slot].p, code_width, rows, (ptrdiff_t)M, 1, yout[slot].as<float>(),
                                        (ptrdiff_t)d, 1, flag.as<int>(), st));
        float *odst = out + (ptrdiff_t)r0 * ors;
        if (out_direct && ors == (ptrdiff_t)d) {
            RB_CUDA_TRY(cudaMemcpyAsync(odst, yout[slot].p, rows * d * sizeof(float), cudaMemcpyDeviceToHost, st));
        } else if (out_direct) {
            RB_CUDA_TRY(cudaMemcpy2DAsync(odst, (size_t)ors * sizeof(float), yout[slot].p, d * sizeof(float), d * sizeof(float),
                                          rows, cudaMemcpyDeviceToHost, st));
        } else {
            RB_CUDA_TRY(cudaMemcpyAsync(sout[slot], yout[slot].p, rows * d * sizeof(float), cudaMemcpyDeviceToHost, st));
            pending_r0[slot] = r0;
            pending_rows[slot] = rows;
        }
    }
    for (int i = 0; i < 2; i++) RB_TRY(drain(i));
    int bad = 0;
    RB_CUDA_TRY(cudaMemcpy(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) return fail(RB_ERR_CODE_RANGE, "a code is >= the number of centroids (%zu)", pq->k);
    return RB_OK;
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *rb_last_error_message(void) { return t_err; }
int rb_abi_version(void) { return RB_ABI_VERSION; }
uint64_t rb_kernel_launch_count(void) { return g_launches.load(); }

rb_status rb_set_encode_algo(int algo)
{
    if (algo != RB_ENCODE_AUTO && algo != RB_ENCODE_EXACT && algo != RB_ENCODE_TENSOR)
        return fail(RB_ERR_INVALID, "unknown encode algo %d", algo);
    g_encode_algo.store(algo);
    return RB_OK;
}

rb_status rb_release_scratch(void)
{
    int dev = 0;
    RB_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (dev >= 0 && dev < 64 && g_pools[dev]) RB_CUDA_TRY(cudaMemPoolTrimTo(g_pools[dev], 0));
    return RB_OK;
}

rb_status rb_set_project_algo(int algo)
{
    if (algo != RB_PROJECT_AUTO && algo != RB_PROJECT_EXACT && algo != RB_PROJECT_TENSOR)
        return fail(RB_ERR_INVALID, "unknown projection algo %d", algo);
    g_project_algo.store(algo);
    return RB_OK;
}

rb_status rb_set_kmeans_update(int ordered)
{
    if (ordered < 0 || ordered > 3) return fail(RB_ERR_INVALID, "kmeans update mode must be 0, 1, 2 or 3");
    g_kmeans_ordered.store(ordered);
    return RB_OK;
}

rb_status rb_pq_create(const float *quantizers, size_t M, size_t k, size_t dsub, const float *projection,
                       rb_pq **out)
{
    if (!out) return fail(RB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!quantizers) return fail(RB_ERR_INVALID, "quantizers is NULL");
    if (M == 0 || k == 0 || dsub == 0)  // pq.rs:39-42
        return fail(RB_ERR_SHAPE, "Attempted to construct a product quantizer without quantizers.");
    RB_TRY(require_device());
    rb_pq *pq = new rb_pq();
    pq->M = M; pq->k = k; pq->dsub = dsub; pq->d = M * dsub;
    cudaGetDevice(&pq->device);
    const size_t qn = M * k * dsub;
    pq->q_host.assign(quantizers, quantizers + qn);
    rb_status st = RB_OK;
    auto body = [&]() -> rb_status {
        RB_CUDA_TRY(cudaMalloc(&pq->q_dev, qn * sizeof(float)));
        RB_CUDA_TRY(cudaMalloc(&pq->cs_dev, M * k * sizeof(float)));
        RB_CUDA_TRY(cudaMemcpy(pq->q_dev, quantizers, qn * sizeof(float), cudaMemcpyHostToDevice));
        RB_TRY(launch_centroid_norms(pq->q_dev, M * k, dsub, pq->cs_dev, nullptr));
        if (projection) {
            pq->proj_host.assign(projection, projection + pq->d * pq->d);
            RB_CUDA_TRY(cudaMalloc(&pq->proj_dev, pq->d * pq->d * sizeof(float)));
            RB_CUDA_TRY(cudaMemcpy(pq->proj_dev, projection, pq->d * pq->d * sizeof(float), cudaMemcpyHostToDevice));
            std::vector<float> rt(pq->d * pq->d);
            for (size_t i = 0; i < pq->d; i++)
                for (size_t j = 0; j < pq->d; j++) rt[j * pq->d + i] = projection[i * pq->d + j];
            RB_CUDA_TRY(cudaMalloc(&pq->projt_dev, pq->d * pq->d * sizeof(float)));
            RB_CUDA_TRY(cudaMemcpy(pq->projt_dev, rt.data(), pq->d * pq->d * sizeof(float), cudaMemcpyHostToDevice));
            RB_TRY(pq->ptc_enc.prepare(pq->proj_dev, projection, pq->d, nullptr));
            RB_TRY(pq->ptc_dec.prepare(pq->projt_dev, rt.data(), pq->d, nullptr));
            float amax = 0.f;
            for (size_t i = 0; i < qn; i++) {
                const float a = std::fabs(quantizers[i]);
                if (!(a <= 3.0e38f)) {  // NaN / Inf
                    amax = -1.f;
                    break;
                }
                amax = a > amax ? a : amax;
            }
            pq->q_absmax = amax;
        }
        RB_TRY(pq->tc.prepare(pq->cb(), nullptr));
        RB_CUDA_TRY(cudaStreamSynchronize(nullptr));
        return RB_OK;
    };
    st = body();
    if (st != RB_OK) {
        rb_pq_destroy(pq);
        return st;
    }
    *out = pq;
    return RB_OK;
}

rb_status rb_pq_create_multi(const float *quantizers, size_t M, size_t k, size_t dsub, const float *projection,
                             const int *devices, int n_devices, rb_pq **out)
{
    if (!out) return fail(RB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices <= 0) return fail(RB_ERR_INVALID, "no devices");
    RB_TRY(require_device());
    int count = 0, prev = 0;
    RB_CUDA_TRY(cudaGetDeviceCount(&count));
    for (int i = 0; i < n_devices; i++)
        if (devices[i] < 0 || devices[i] >= count) return fail(RB_ERR_INVALID, "device %d is not visible (%d devices)", devices[i], count);
    cudaGetDevice(&prev);
    std::vector<rb_pq *> reps;
    rb_status st = RB_OK;
    for (int i = 0; i < n_devices && st == RB_OK; i++) {
        rb_pq *r = nullptr;
        if (cudaSetDevice(devices[i]) != cudaSuccess) st = fail(RB_ERR_CUDA, "cudaSetDevice(%d) failed", devices[i]);
        if (st == RB_OK) st = rb_pq_create(quantizers, M, k, dsub, projection, &r);
        if (r) reps.push_back(r);
    }
    cudaSetDevice(prev);
    if (st != RB_OK) {
        for (rb_pq *r : reps) rb_pq_destroy(r);
        return st;
    }
    reps[0]->peers.assign(reps.begin() + 1, reps.end());
    *out = reps[0];
    return RB_OK;
}

int rb_pq_n_devices(const rb_pq *pq) { return pq ? 1 + (int)pq->peers.size() : 0; }

rb_status rb_set_host_copy_threads(int n_threads)
{
    if (n_threads < 1 || n_threads > 64) return fail(RB_ERR_INVALID, "n_threads must be in [1, 64]");
    g_copy_threads.store(n_threads);
    return RB_OK;
}

void rb_pq_destroy(rb_pq *pq)
{
    if (!pq) return;
    if (!pq->peers.empty()) {
        int prev = 0;
        cudaGetDevice(&prev);
        for (rb_pq *r : pq->peers) rb_pq_destroy(r);
        pq->peers.clear();
        cudaSetDevice(pq->device);
        pq->tc.release();
        pq->ptc_enc.release();
        pq->ptc_dec.release();
        cudaFree(pq->q_dev);
        cudaFree(pq->cs_dev);
        cudaFree(pq->proj_dev);
        cudaFree(pq->projt_dev);
        cudaSetDevice(prev);
        delete pq;
        return;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != pq->device) cudaSetDevice(pq->device);
    pq->tc.release();
    pq->ptc_enc.release();
    pq->ptc_dec.release();
    cudaFree(pq->q_dev);
    cudaFree(pq->cs_dev);
    cudaFree(pq->proj_dev);
    cudaFree(pq->projt_dev);
    if (cur != pq->device) cudaSetDevice(cur);
    delete pq;
}

size_t rb_pq_quantized_len(const rb_pq *pq) { return pq ? pq->M : 0; }
size_t rb_pq_reconstructed_len(const rb_pq *pq) { return pq ? pq->d : 0; }
size_t rb_pq_n_quantizer_centroids(const rb_pq *pq) { return pq ? pq->k : 0; }
int rb_pq_has_projection(const rb_pq *pq) { return pq && pq->proj_dev ? 1 : 0; }

rb_status rb_pq_subquantizers(const rb_pq *pq, float *out_host)
{
    if (!pq || !out_host) return fail(RB_ERR_INVALID, "NULL argument");
    memcpy(out_host, pq->q_host.data(), pq->q_host.size() * sizeof(float));
    return RB_OK;
}

rb_status rb_pq_projection(const rb_pq *pq, float *out_host)
{
    if (!pq || !out_host) return fail(RB_ERR_INVALID, "NULL argument");
    if (pq->proj_host.empty()) return fail(RB_ERR_INVALID, "quantizer has no projection");
    memcpy(out_host, pq->proj_host.data(), pq->proj_host.size() * sizeof(float));
    return RB_OK;
}

rb_status rb_pq_quantize_batch(const rb_pq *pq, const float *x, size_t n, ptrdiff_t rs, ptrdiff_t cs, void *codes,
                               int code_width, ptrdiff_t crs, ptrdiff_t ccs, int mem_kind, void *stream)
{
    if (!pq) return fail(RB_ERR_INVALID, "pq is NULL");
    if (!valid_width(code_width)) return fail(RB_ERR_INVALID, "code_width must be 1, 2, 4 or 8");
    if (mem_kind != RB_MEM_HOST && mem_kind != RB_MEM_DEVICE) return fail(RB_ERR_INVALID, "bad mem_kind");
    if (n == 0) return RB_OK;
    if (!x || !codes) return fail(RB_ERR_INVALID, "NULL data pointer");
    RB_TRY(require_device());
    if (mem_kind == RB_MEM_DEVICE)
        return quantize_batch_device(pq, x, n, rs, cs, codes, code_width, crs, ccs, (cudaStream_t)stream);
    // host memory: contiguous row blocks over the quantizer's devices (one device unless rb_pq_create_multi made it)
    return over_devices(pq, n, [&](const rb_pq *rep, size_t r0, size_t rows) {
        return quantize_batch_host(rep, x + (ptrdiff_t)r0 * rs, rows, rs, cs,
                                   reinterpret_cast<char *>(codes) + (ptrdiff_t)r0 * crs * code_width, code_width, crs, ccs);
    });
}

rb_status rb_pq_quantize_vector(const rb_pq *pq, const float *x, ptrdiff_t sx, void *codes, int code_width,
                                ptrdiff_t cstride, int mem_kind, void *stream)
{
    if (!pq || !x || !codes) return fail(RB_ERR_INVALID, "NULL argument");
    if (!valid_width(code_width)) return fail(RB_ERR_INVALID, "code_width must be 1, 2, 4 or 8");
    if (code_width < 8) {  // primitives.rs:31-34
        const uint64_t maxv = code_width == 1 ? 0xffull : code_width == 2 ? 0xffffull : 0xffffffffull;
        if ((uint64_t)(pq->k - 1) > maxv)
            return fail(RB_ERR_CODE_TYPE, "Cannot store centroids in quantizer index type");
    }
    RB_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = pq->d, M = pq->M;
    Workspace scratch;
    RB_TRY(scratch.alloc(d * sizeof(float), st));
    if (mem_kind == RB_MEM_DEVICE)
        return launch_quantize_vector(pq->cb(), pq->proj_dev, x, sx, codes, code_width, cstride, scratch.as<float>(), st);
    Workspace xd, cd;
    RB_TRY(xd.alloc(d * sizeof(float), st));
    RB_TRY(cd.alloc(M * code_width, st));
    std::vector<float> xh(d);
    for (size_t i = 0; i < d; i++) xh[i] = x[(ptrdiff_t)i * sx];
    RB_CUDA_TRY(cudaMemcpyAsync(xd.p, xh.data(), d * sizeof(float), cudaMemcpyHostToDevice, st));
    // the packed copy is contiguous; a strided host view must still take the sequential-dot arithmetic:
    // pass a fake stride of 2 over a doubled buffer?  No — keep it simple and exact: re-expand on device.
    Workspace xexp;
    const float *xin = xd.as<float>();
    ptrdiff_t xin_stride = 1;
    if (sx != 1 && !pq->proj_dev && d > 1) {
        RB_TRY(xexp.alloc(2 * d * sizeof(float), st));
        RB_CUDA_TRY(cudaMemcpy2DAsync(xexp.p, 2 * sizeof(float), xd.p, sizeof(float), sizeof(float), d,
                                      cudaMemcpyDeviceToDevice, st));
        xin = xexp.as<float>();
        xin_stride = 2;
    }
    RB_TRY(launch_quantize_vector(pq->cb(), pq->proj_dev, xin, xin_stride, cd.p, code_width, 1, scratch.as<float>(), st));
    std::vector<unsigned char> ch(M * code_width);
    RB_CUDA_TRY(cudaMemcpyAsync(ch.data(), cd.p, M * code_width, cudaMemcpyDeviceToHost, st));
    RB_CUDA_TRY(cudaStreamSynchronize(st));
    host_unpack(ch.data(), M, 1, codes, cstride, 1, (size_t)code_width);
    return RB_OK;
}

rb_status rb_pq_reconstruct_batch(const rb_pq *pq, const void *codes, int code_width, size_t n, ptrdiff_t crs,
                                  ptrdiff_t ccs, float *out, ptrdiff_t ors, ptrdiff_t ocs, int mem_kind, void *stream)
{
    if (!pq) return fail(RB_ERR_INVALID, "pq is NULL");
    if (!valid_width(code_width)) return fail(RB_ERR_INVALID, "code_width must be 1, 2, 4 or 8");
    if (mem_kind != RB_MEM_HOST && mem_kind != RB_MEM_DEVICE) return fail(RB_ERR_INVALID, "bad mem_kind");
    if (n == 0) return RB_OK;
    if (!codes || !out) return fail(RB_ERR_INVALID, "NULL data pointer");
    RB_TRY(require_device());
    if (mem_kind == RB_MEM_HOST)
        return over_devices(pq, n, [&](const rb_pq *rep, size_t r0, size_t rows) {
            return reconstruct_batch_host(rep, reinterpret_cast<const char *>(codes) + (ptrdiff_t)r0 * crs * code_width, code_width,
                                          rows, crs, ccs, out + (ptrdiff_t)r0 * ors, ors, ocs);
        });
    cudaStream_t st = (cudaStream_t)stream;
    Workspace flag;
    RB_TRY(flag.alloc(sizeof(int), st));
    RB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    RB_TRY(reconstruct_batch_device(pq, codes, code_width, n, crs, ccs, out, ors, ocs, flag.as<int>(), st));
    // reporting an out-of-range code needs the result, so the call synchronises `stream` -- unless no code value
    // can be out of range (u8 codes of a 256-centroid quantizer), in which case it stays asynchronous
    if (code_width == 1 && pq->k >= 256) return RB_OK;
    int bad = 0;
    RB_CUDA_TRY(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    RB_CUDA_TRY(cudaStreamSynchronize(st));
    if (bad) return fail(RB_ERR_CODE_RANGE, "a code is >= the number of centroids (%zu)", pq->k);
    return RB_OK;
}

rb_status rb_pq_reconstruct(const rb_pq *pq, const void *codes, int code_width, ptrdiff_t cstride, float *out,
                            ptrdiff_t ostride, int mem_kind, void *stream)
{
    if (!pq || !codes || !out) return fail(RB_ERR_INVALID, "NULL argument");
    if (!valid_width(code_width)) return fail(RB_ERR_INVALID, "code_width must be 1, 2, 4 or 8");
    RB_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = pq->d, M = pq->M;
    Workspace scratch, flag;
    RB_TRY(scratch.alloc(d * sizeof(float), st));
    RB_TRY(flag.alloc(sizeof(int), st));
    RB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    int bad = 0;
    if (mem_kind == RB_MEM_DEVICE) {
        RB_TRY(launch_reconstruct_vector(pq->cb(), pq->proj_dev, codes, code_width, cstride, out, ostride,
                                         scratch.as<float>(), flag.as<int>(), st));
    } else {
        Workspace cd, od;
        RB_TRY(cd.alloc(M * code_width, st));
        RB_TRY(od.alloc(d * sizeof(float), st));
        std::vector<unsigned char> ch(M * code_width);
        host_pack(codes, M, 1, cstride, 1, (size_t)code_width, ch.data());
        RB_CUDA_TRY(cudaMemcpyAsync(cd.p, ch.data(), M * code_width, cudaMemcpyHostToDevice, st));
        RB_TRY(launch_reconstruct_vector(pq->cb(), pq->proj_dev, cd.p, code_width, 1, od.as<float>(), 1,
                                         scratch.as<float>(), flag.as<int>(), st));
        std::vector<float> oh(d);
        RB_CUDA_TRY(cudaMemcpyAsync(oh.data(), od.p, d * sizeof(float), cudaMemcpyDeviceToHost, st));
        RB_CUDA_TRY(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        RB_CUDA_TRY(cudaStreamSynchronize(st));
        if (!bad)
            for (size_t i = 0; i < d; i++) out[(ptrdiff_t)i * ostride] = oh[i];
    }
    if (mem_kind == RB_MEM_DEVICE) {
        RB_CUDA_TRY(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        RB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (bad) return fail(RB_ERR_CODE_RANGE, "a code is >= the number of centroids (%zu)", pq->k);
    return RB_OK;
}

rb_status rb_check_quantizer_invariants(size_t n_subquantizers, uint32_t n_bits, size_t n_iterations,
                                        size_t n_attempts, size_t n_rows, size_t n_cols, uint64_t *detail)
{
    if (n_subquantizers == 0 || n_subquantizers > n_cols) {  // pq.rs:70-75
        if (detail) *detail = n_cols;
        return fail(RB_ERR_N_SUBQUANTIZERS_RANGE, "The number of subquantizers must be between 1 and %zu, was %zu",
                    n_cols, n_subquantizers);
    }
    // pq.rs:77: (nrows as f64).log2().trunc() as u32  (saturating cast: 0 rows -> 0)
    uint32_t max_bits = 0;
    for (size_t v = n_rows; v > 1; v >>= 1) max_bits++;
    if (n_bits == 0 || n_bits > max_bits) {  // pq.rs:78-82
        if (detail) *detail = max_bits;
        return fail(RB_ERR_N_SUBQUANTIZER_BITS, "The number of subquantizers bits must be between 1 and %u", max_bits);
    }
    if (n_cols % n_subquantizers != 0)  // pq.rs:84-89
        return fail(RB_ERR_NUMBER_SUBQUANTIZERS,
                    "The number of columns (%zu) is not exactly dividable by the number of subquantizers (%zu)",
                    n_cols, n_subquantizers);
    if (n_iterations == 0)  // pq.rs:91-93
        return fail(RB_ERR_N_ITERATIONS, "The number of quantization iterations must be >= 1");
    if (n_attempts == 0)  // pq.rs:95-97
        return fail(RB_ERR_N_ATTEMPTS, "The number of quantization attempts per iteration must be >= 1");
    return RB_OK;
}

size_t rb_kmeans_packed_len(size_t M, size_t k, size_t dsub) { return M * k * dsub + M * k + M; }

size_t rb_kmeans_code_pitch(size_t n_local) { return (n_local + 15) / 16 * 16 + 16; }
int rb_kmeans_code_width(size_t k) { return k <= 256 ? 1 : 4; }

rb_status rb_kmeans_assign(const float *x, size_t n_local, ptrdiff_t ldx, const float *centroids, size_t M, size_t k,
                           size_t dsub, void *codes, void *stream)
{
    return rb::kmeans_assign_strided(x, n_local, ldx, centroids, M, k, dsub, codes, (ptrdiff_t)rb_kmeans_code_pitch(n_local),
                                     (cudaStream_t)stream);
}

}  // extern "C"

// column-major codes with an explicit column stride (elements): dist.cu points it at the peer-mapped code matrix
rb_status rb::kmeans_assign_strided(const float *x, size_t n_local, ptrdiff_t ldx, const float *centroids, size_t M, size_t k,
                                    size_t dsub, void *codes, ptrdiff_t col_stride, cudaStream_t st)
{
    if (!centroids || !codes || (n_local && !x)) return fail(RB_ERR_INVALID, "NULL argument");
    if (M == 0 || k == 0 || dsub == 0) return fail(RB_ERR_SHAPE, "Cannot cluster instances with zero centroids.");
    RB_TRY(require_device());
    Workspace cs;
    RB_TRY(cs.alloc(M * k * sizeof(float), st));
    RB_TRY(launch_centroid_norms(centroids, M * k, dsub, cs.as<float>(), st));
    const DeviceCodebook cb{centroids, cs.as<float>(), M, k, dsub};
    TensorOperands tc;
    if (g_encode_algo.load() != RB_ENCODE_EXACT) RB_TRY(tc.prepare(cb, st));
    // column-major [M][pitch]: the tensor kernel's row-per-thread stores and the per-subquantizer sort both touch
    // contiguous bytes
    const rb_status s = encode_device(cb, &tc, x, n_local, ldx, 0, codes, rb_kmeans_code_width(k), 1, col_stride, st);  // kmeans.rs:319
    tc.release_async(st);
    return s;
}

extern "C" {

rb_status rb_kmeans_accumulate(const float *x, size_t n_local, ptrdiff_t ldx, const void *codes, size_t M, size_t k,
                               size_t dsub, const float *packed_before, float *packed, void *stream)
{
    if (!codes || !packed || (n_local && !x)) return fail(RB_ERR_INVALID, "NULL argument");
    if (M == 0 || k == 0 || dsub == 0) return fail(RB_ERR_SHAPE, "Cannot cluster instances with zero centroids.");
    RB_TRY(require_device());
    const int width = rb_kmeans_code_width(k);
    return launch_kmeans_accumulate(x, n_local, ldx, width == 1 ? reinterpret_cast<const uint8_t *>(codes) : nullptr,
                                    width == 4 ? reinterpret_cast<const uint32_t *>(codes) : nullptr,
                                    rb_kmeans_code_pitch(n_local), M, k, dsub, packed_before, packed,
                                    g_kmeans_ordered.load(), (cudaStream_t)stream);
}

rb_status rb_kmeans_assign_accumulate_from(const float *x, size_t n_local, ptrdiff_t ldx, const float *centroids,
                                           size_t M, size_t k, size_t dsub, const float *packed_before,
                                           float *packed, void *stream)
{
    if (!centroids || !packed || (n_local && !x)) return fail(RB_ERR_INVALID, "NULL argument");
    if (M == 0 || k == 0 || dsub == 0) return fail(RB_ERR_SHAPE, "Cannot cluster instances with zero centroids.");
    RB_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    Workspace codes;  // assignments are an internal temporary here
    RB_TRY(codes.alloc(M * rb_kmeans_code_pitch(n_local) * (size_t)rb_kmeans_code_width(k), st));
    RB_TRY(rb_kmeans_assign(x, n_local, ldx, centroids, M, k, dsub, codes.p, stream));
    return rb_kmeans_accumulate(x, n_local, ldx, codes.p, M, k, dsub, packed_before, packed, stream);
}

rb_status rb_kmeans_assign_accumulate(const float *x, size_t n_local, ptrdiff_t ldx, const float *centroids, size_t M,
                                      size_t k, size_t dsub, float *packed, void *stream)
{
    return rb_kmeans_assign_accumulate_from(x, n_local, ldx, centroids, M, k, dsub, nullptr, packed, stream);
}

rb_status rb_kmeans_finalize(const float *packed, size_t M, size_t k, size_t dsub, uint64_t n_total, float *centroids,
                             float *loss_or_null, void *stream)
{
    if (!packed || !centroids) return fail(RB_ERR_INVALID, "NULL argument");
    RB_TRY(require_device());
    return launch_kmeans_finalize(packed, M, k, dsub, n_total, centroids, loss_or_null, (cudaStream_t)stream);
}

rb_status rb_pq_train(const float *instances, size_t n, size_t d, ptrdiff_t rs, ptrdiff_t cs, size_t M,
                      uint32_t n_bits, size_t n_iterations, size_t n_attempts, const float *initial, float *loss_out,
                      int mem_kind, void *stream, rb_pq **out)
{
    if (!out) return fail(RB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    RB_TRY(rb_check_quantizer_invariants(M, n_bits, n_iterations, n_attempts, n, d, nullptr));  // pq.rs:213-219
    if (!instances || !initial) return fail(RB_ERR_INVALID, "NULL argument");
    const size_t k = (size_t)1 << n_bits;  // pq.rs:233
    if (k >= n)  // kmeans.rs:62-67 (RandomInstanceCentroids asserts k < n)
        return fail(RB_ERR_K_MEANS_K, "Cannot pick more centroids than instances: %zu instances, %zu centroids", n, k);
    RB_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const size_t dsub = d / M, qn = M * k * dsub;

    // instances resident and row-major on the device for the whole run
    Workspace xbuf;
    const float *x = instances;
    ptrdiff_t ldx = rs;
    if (mem_kind == RB_MEM_HOST) {
        RB_TRY(xbuf.alloc(n * d * sizeof(float), st));
        if (cs == 1 && rs == (ptrdiff_t)d) {
            RB_CUDA_TRY(cudaMemcpyAsync(xbuf.p, instances, n * d * sizeof(float), cudaMemcpyHostToDevice, st));
        } else if (cs == 1 && rs >= (ptrdiff_t)d) {
            RB_CUDA_TRY(cudaMemcpy2DAsync(xbuf.p, d * sizeof(float), instances, (size_t)rs * sizeof(float),
                                          d * sizeof(float), n, cudaMemcpyHostToDevice, st));
        } else {
            std::vector<float> tmp(n * d);
            host_pack(instances, n, d, rs, cs, sizeof(float), tmp.data());
            RB_CUDA_TRY(cudaMemcpyAsync(xbuf.p, tmp.data(), n * d * sizeof(float), cudaMemcpyHostToDevice, st));
            RB_CUDA_TRY(cudaStreamSynchronize(st));
        }
        x = xbuf.as<float>();
        ldx = (ptrdiff_t)d;
    } else if (cs != 1) {
        RB_TRY(xbuf.alloc(n * d * sizeof(float), st));
        RB_TRY(launch_pack_rows(instances, n, d, rs, cs, xbuf.as<float>(), st));
        x = xbuf.as<float>();
        ldx = (ptrdiff_t)d;
    }

    // The loop runs on the single-rank form of the sharded k-means state (csrc/dist.cu): it keeps sum ||x||^2 per
    // subquantizer in FP64 (fixed order, computed once: the instances never change) so that the losses that rank the
    // attempts (pq.rs:183-187) are run-to-run identical, and the subquantizer-major slabs the streaming update reads.
    rb_comm *self = nullptr;
    rb_kmeans_dist *state = nullptr;
    RB_TRY(rb_comm_create(nullptr, 0, 1, &self));
    struct Cleanup {
        rb_comm *&c;
        rb_kmeans_dist *&s;
        ~Cleanup()
        {
            rb_kmeans_dist_destroy(s);
            rb_comm_destroy(c);
        }
    } cleanup{self, state};
    RB_TRY(rb_kmeans_dist_create(self, x, n, ldx, M, k, dsub, stream, &state));
    Workspace cen, loss_dev;
    RB_TRY(cen.alloc(qn * sizeof(float), st));
    RB_TRY(loss_dev.alloc(M * sizeof(float), st));
    std::vector<float> best_q(qn), cand_q(qn), best_loss(M, 0.f), cand_loss(M, 0.f);
    for (size_t a = 0; a < n_attempts; a++) {  // pq.rs:168-183, all subquantizers advance together
        RB_CUDA_TRY(cudaMemcpyAsync(cen.p, initial + a * qn, qn * sizeof(float), cudaMemcpyHostToDevice, st));
        for (size_t it = 0; it < n_iterations; it++)  // kmeans.rs:279-284
            RB_TRY(rb_kmeans_dist_iterate(state, cen.as<float>(), it + 1 == n_iterations ? loss_dev.as<float>() : nullptr, stream));
        RB_CUDA_TRY(cudaMemcpyAsync(cand_q.data(), cen.p, qn * sizeof(float), cudaMemcpyDeviceToHost, st));
        RB_CUDA_TRY(cudaMemcpyAsync(cand_loss.data(), loss_dev.p, M * sizeof(float), cudaMemcpyDeviceToHost, st));
        RB_CUDA_TRY(cudaStreamSynchronize(st));
        for (size_t m = 0; m < M; m++) {  // pq.rs:184-187: min_by_key(OrderedFloat(loss)) keeps the first minimum
            const float l = cand_loss[m], b = best_loss[m];
            const bool less = (l < b) || ((b != b) && (l == l));
            if (a == 0 || less) {
                best_loss[m] = l;
                memcpy(best_q.data() + m * k * dsub, cand_q.data() + m * k * dsub, k * dsub * sizeof(float));
            }
        }
    }
    if (loss_out) memcpy(loss_out, best_loss.data(), M * sizeof(float));
    return rb_pq_create(best_q.data(), M, k, dsub, nullptr, out);
}

// ---- Opq training --------------------------------------------------------------------------------------------------
rb_status rb_set_gram_algo(int algo)
{
    if (algo < 0 || algo > 2) return fail(RB_ERR_INVALID, "gram algo must be 0 (auto), 1 (CUDA cores) or 2 (tensor cores)");
    set_gram_algo(algo);
    return RB_OK;
}

rb_status rb_covariance(const float *x, size_t n, size_t d, ptrdiff_t ldx, float *cov_out, void *stream)
{
    if (!x || !cov_out) return fail(RB_ERR_INVALID, "NULL argument");
    if (n == 0) return fail(RB_ERR_SHAPE, "Cannot compute a covariance from zero observations");  // linalg.rs:24-27
    if (d == 0 || ldx < (ptrdiff_t)d) return fail(RB_ERR_SHAPE, "bad shape (d=%zu, row stride %td)", d, ldx);
    RB_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    Workspace means;
    RB_TRY(means.alloc(d * sizeof(float), st));
    RB_TRY(launch_column_means(x, n, d, ldx, means.as<float>(), st));  // linalg.rs:30
    // centered.t().dot(&centered.map(|v| *v / normalization))  linalg.rs:31-41
    return launch_gram(x, ldx, x, ldx, n, d, d, means.as<float>(), means.as<float>(), (float)n - 1.0f, cov_out, st);
}

rb_status rb_opq_train_iteration(const float *x, size_t n, size_t d, ptrdiff_t ldx, const float *projection, float *centroids,
                                 size_t M, size_t k, float *xty_out, void *stream)
{
    if (!x || !projection || !centroids || !xty_out) return fail(RB_ERR_INVALID, "NULL argument");
    if (M == 0 || k == 0 || d == 0 || d % M != 0 || ldx < (ptrdiff_t)d) return fail(RB_ERR_SHAPE, "bad shape");
    if (k >= n) return fail(RB_ERR_K_MEANS_K, "Cannot pick more centroids than instances: %zu instances, %zu centroids", n, k);
    RB_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const size_t dsub = d / M;
    Workspace rx, packed, cs, codes;
    RB_TRY(rx.alloc(n * d * sizeof(float), st));
    RB_TRY(packed.alloc(rb_kmeans_packed_len(M, k, dsub) * sizeof(float), st));
    RB_TRY(launch_project(x, n, d, ldx, 1, projection, 0, rx.as<float>(), st));  // opq.rs:173
    // one k-means step per subquantizer (opq.rs:174,191-209): kmeans.rs:319-325, all M at once
    RB_TRY(rb_kmeans_assign_accumulate(rx.as<float>(), n, (ptrdiff_t)d, centroids, M, k, dsub, packed.as<float>(), stream));
    RB_TRY(launch_kmeans_finalize(packed.as<float>(), M, k, dsub, n, centroids, nullptr, st));
    // quantize -> reconstruct with the NEW centroids (opq.rs:180-182); the reconstruction reuses rx's buffer as the
    // reference does
    const int width = rb_kmeans_code_width(k);
    RB_TRY(codes.alloc(n * M * (size_t)width, st));
    RB_TRY(cs.alloc(M * k * sizeof(float), st));
    RB_TRY(launch_centroid_norms(centroids, M * k, dsub, cs.as<float>(), st));
    const DeviceCodebook cb{centroids, cs.as<float>(), M, k, dsub};
    TensorOperands tc;
    if (g_encode_algo.load() != RB_ENCODE_EXACT) RB_TRY(tc.prepare(cb, st));
    rb_status s = encode_device(cb, &tc, rx.as<float>(), n, (ptrdiff_t)d, 0, codes.p, width, (ptrdiff_t)M, 1, st);
    tc.release_async(st);
    RB_TRY(s);
    Workspace flag;
    RB_TRY(flag.alloc(sizeof(int), st));
    RB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    RB_TRY(launch_gather(cb, codes.p, width, n, (ptrdiff_t)M, 1, rx.as<float>(), (ptrdiff_t)d, flag.as<int>(), st));
    // instances.t().dot(&reconstructed)  opq.rs:187
    return launch_gram(x, ldx, rx.as<float>(), (ptrdiff_t)d, n, d, d, nullptr, nullptr, 1.0f, xty_out, st);
}

// ---- A = f64: defined answers, no silent down-conversion (include/reductive_b200.h) -------------------------------
static rb_status f64_unsupported(const char *what)
{
    return fail(RB_ERR_UNSUPPORTED,
                "%s: this library's kernels are f32 only (Pq<f64> stays on the reference's CPU path; f32 arithmetic cannot "
                "reproduce f64 codes bit for bit)", what);
}

rb_status rb_pq_create_f64(const double *, size_t, size_t, size_t, const double *, rb_pq **out)
{
    if (out) *out = nullptr;
    return f64_unsupported("rb_pq_create_f64");
}

rb_status rb_pq_quantize_batch_f64(const rb_pq *, const double *, size_t, ptrdiff_t, ptrdiff_t, void *, int, ptrdiff_t, ptrdiff_t,
                                   int, void *)
{
    return f64_unsupported("rb_pq_quantize_batch_f64");
}

rb_status rb_pq_reconstruct_batch_f64(const rb_pq *, const void *, int, size_t, ptrdiff_t, ptrdiff_t, double *, ptrdiff_t,
                                      ptrdiff_t, int, void *)
{
    return f64_unsupported("rb_pq_reconstruct_batch_f64");
}

rb_status rb_pq_train_f64(const double *, size_t, size_t, ptrdiff_t, ptrdiff_t, size_t, uint32_t, size_t, size_t, const double *,
                          double *, int, void *, rb_pq **out)
{
    if (out) *out = nullptr;
    return f64_unsupported("rb_pq_train_f64");
}

rb_status rb_project_rows(const float *x, size_t n, size_t d, ptrdiff_t rs, ptrdiff_t cs, const float *r_dev,
                          int transpose_r, float *out, void *stream)
{
    if ((n && (!x || !out)) || !r_dev) return fail(RB_ERR_INVALID, "NULL argument");
    RB_TRY(require_device());
    return launch_project(x, n, d, rs, cs, r_dev, transpose_r, out, (cudaStream_t)stream);
}

// ---- caller-side quantized storage (kernels in qstore.cu) ------------------------------------------
struct rb_qstore {
    const rb_pq *pq = nullptr;  // borrowed
    size_t n = 0;
    uint8_t *codes = nullptr;  // DEVICE [n][M], allocation padded to 16 bytes
    float *norms = nullptr;    // DEVICE [n] or null
};

namespace {
struct DeviceScope {  // make the quantizer's device current for a host-memory call
    int prev = 0;
    bool switched = false;
    rb_status enter(int device)
    {
        RB_CUDA_TRY(cudaGetDevice(&prev));
        if (prev != device) {
            RB_CUDA_TRY(cudaSetDevice(device));
            switched = true;
        }
        return RB_OK;
    }
    ~DeviceScope()
    {
        if (switched) cudaSetDevice(prev);
    }
};
}  // namespace

void rb_qstore_destroy(rb_qstore *s)
{
    if (!s) return;
    DeviceScope scope;
    if (s->pq) scope.enter(s->pq->device);
    cudaFree(s->codes);
    cudaFree(s->norms);
    delete s;
}

rb_status rb_qstore_create(const rb_pq *pq, const uint8_t *codes, size_t n, ptrdiff_t code_row_stride,
                           const float *norms_or_null, int mem_kind, void *stream, rb_qstore **out)
{
    if (!out) return fail(RB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!pq || (n && !codes)) return fail(RB_ERR_INVALID, "NULL argument");
    if (mem_kind != RB_MEM_HOST && mem_kind != RB_MEM_DEVICE) return fail(RB_ERR_INVALID, "bad mem_kind");
    if (pq->k > 256) return fail(RB_ERR_CODE_TYPE, "Cannot store centroids in quantizer index type");  // u8 codes
    if (code_row_stride < (ptrdiff_t)pq->M) return fail(RB_ERR_SHAPE, "code rows overlap (stride %td < %zu)", code_row_stride, pq->M);
    RB_TRY(require_device());
    DeviceScope scope;
    if (mem_kind == RB_MEM_HOST) RB_TRY(scope.enter(pq->device));
    cudaStream_t st = (cudaStream_t)stream;
    rb_qstore *s = new rb_qstore();
    s->pq = pq;
    s->n = n;
    auto body = [&]() -> rb_status {
        const size_t M = pq->M, bytes = (n * M + 15) & ~(size_t)15;
        RB_CUDA_TRY(cudaMalloc((void **)&s->codes, bytes ? bytes : 16));
        RB_CUDA_TRY(cudaMemsetAsync(s->codes, 0, bytes ? bytes : 16, st));
        const cudaMemcpyKind kind = mem_kind == RB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        if (n) RB_CUDA_TRY(cudaMemcpy2DAsync(s->codes, M, codes, (size_t)code_row_stride, M, n, kind, st));
        if (norms_or_null) {
            RB_CUDA_TRY(cudaMalloc((void **)&s->norms, (n ? n : 1) * sizeof(float)));
            if (n) RB_CUDA_TRY(cudaMemcpyAsync(s->norms, norms_or_null, n * sizeof(float), kind, st));
        }
        // every stored code must name a centroid (quantize_batch guarantees it; reconstruct would panic otherwise)
        Workspace flag;
        RB_TRY(flag.alloc(sizeof(int), st));
        RB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
        RB_TRY(launch_qstore_check_codes(s->codes, n * M, pq->k, flag.as<int>(), st));
        int bad = 0;
        RB_CUDA_TRY(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        RB_CUDA_TRY(cudaStreamSynchronize(st));
        if (bad) return fail(RB_ERR_CODE_RANGE, "a code is >= the number of centroids (%zu)", pq->k);
        return RB_OK;
    };
    const rb_status r = body();
    if (r != RB_OK) {
        rb_qstore_destroy(s);
        return r;
    }
    *out = s;
    return RB_OK;
}

size_t rb_qstore_len(const rb_qstore *s) { return s ? s->n : 0; }
int rb_qstore_has_norms(const rb_qstore *s) { return s && s->norms ? 1 : 0; }

rb_status rb_qstore_embeddings(const rb_qstore *s, const uint64_t *indices, size_t n_idx, float *out, ptrdiff_t ors,
                               ptrdiff_t ocs, int mem_kind, void *stream)
{
    if (!s) return fail(RB_ERR_INVALID, "store is NULL");
    if (mem_kind != RB_MEM_HOST && mem_kind != RB_MEM_DEVICE) return fail(RB_ERR_INVALID, "bad mem_kind");
    if (n_idx == 0) return RB_OK;
    if (!indices || !out) return fail(RB_ERR_INVALID, "NULL data pointer");
    RB_TRY(require_device());
    const rb_pq *pq = s->pq;
    const size_t M = pq->M, d = pq->d;
    DeviceScope scope;
    if (mem_kind == RB_MEM_HOST) RB_TRY(scope.enter(pq->device));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace idx_dev, sel, nsel, flag, dense;
    const unsigned long long *idx = reinterpret_cast<const unsigned long long *>(indices);
    if (mem_kind == RB_MEM_HOST) {
        RB_TRY(idx_dev.alloc(n_idx * sizeof(uint64_t), st));
        RB_CUDA_TRY(cudaMemcpyAsync(idx_dev.p, indices, n_idx * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        idx = idx_dev.as<unsigned long long>();
    }
    RB_TRY(sel.alloc(n_idx * M + 16, st));
    if (s->norms) RB_TRY(nsel.alloc(n_idx * sizeof(float), st));
    RB_TRY(flag.alloc(2 * sizeof(int), st));
    RB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), st));
    RB_TRY(launch_qstore_select(s->codes, s->n, M, idx, n_idx, sel.as<uint8_t>(), s->norms, s->norms ? nsel.as<float>() : nullptr,
                                flag.as<int>(), st));
    float *dst = out;
    ptrdiff_t drs = ors, dcs = ocs;
    if (mem_kind == RB_MEM_HOST) {
        RB_TRY(dense.alloc(n_idx * d * sizeof(float), st));
        dst = dense.as<float>();
        drs = (ptrdiff_t)d;
        dcs = 1;
    }
    // reconstruct (pq.rs:303-347, projection included), then scale by the row's norm
    RB_TRY(reconstruct_batch_device(pq, sel.p, 1, n_idx, (ptrdiff_t)M, 1, dst, drs, dcs, flag.as<int>() + 1, st));
    if (s->norms) RB_TRY(launch_qstore_scale_rows(dst, drs, dcs, n_idx, d, nsel.as<float>(), st));
    int bad = 0;
    RB_CUDA_TRY(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (mem_kind == RB_MEM_HOST) {
        if (ocs == 1)
            RB_CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)ors * sizeof(float), dst, d * sizeof(float), d * sizeof(float), n_idx,
                                          cudaMemcpyDeviceToHost, st));
        else {
            std::vector<float> tmp(n_idx * d);
            RB_CUDA_TRY(cudaMemcpyAsync(tmp.data(), dst, n_idx * d * sizeof(float), cudaMemcpyDeviceToHost, st));
            RB_CUDA_TRY(cudaStreamSynchronize(st));
            for (size_t i = 0; i < n_idx; i++)
                for (size_t c = 0; c < d; c++) out[(ptrdiff_t)i * ors + (ptrdiff_t)c * ocs] = tmp[i * d + c];
        }
    }
    RB_CUDA_TRY(cudaStreamSynchronize(st));
    if (bad) return fail(RB_ERR_INVALID, "an index is >= the number of stored rows (%zu)", s->n);
    return RB_OK;
}

rb_status rb_qstore_dot(const rb_qstore *s, const float *queries, size_t nq, ptrdiff_t qrs, ptrdiff_t qcs, float *out,
                        ptrdiff_t out_row_stride, int mem_kind, void *stream)
{
    if (!s) return fail(RB_ERR_INVALID, "store is NULL");
    if (mem_kind != RB_MEM_HOST && mem_kind != RB_MEM_DEVICE) return fail(RB_ERR_INVALID, "bad mem_kind");
    if (nq == 0 || s->n == 0) return RB_OK;
    if (!queries || !out) return fail(RB_ERR_INVALID, "NULL data pointer");
    if (out_row_stride < (ptrdiff_t)s->n) return fail(RB_ERR_SHAPE, "score rows overlap (stride %td < %zu)", out_row_stride, s->n);
    RB_TRY(require_device());
    const rb_pq *pq = s->pq;
    const size_t M = pq->M, d = pq->d, k = pq->k, n = s->n;
    DeviceScope scope;
    if (mem_kind == RB_MEM_HOST) RB_TRY(scope.enter(pq->device));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace qd, qrot, lut, scores;
    const float *q = queries;
    ptrdiff_t rs = qrs, cs = qcs;
    if (mem_kind == RB_MEM_HOST) {
        std::vector<float> packed(nq * d);
        for (size_t i = 0; i < nq; i++)
            for (size_t c = 0; c < d; c++) packed[i * d + c] = queries[(ptrdiff_t)i * qrs + (ptrdiff_t)c * qcs];
        RB_TRY(qd.alloc(nq * d * sizeof(float), st));
        RB_CUDA_TRY(cudaMemcpyAsync(qd.p, packed.data(), nq * d * sizeof(float), cudaMemcpyHostToDevice, st));
        RB_CUDA_TRY(cudaStreamSynchronize(st));  // `packed` goes away
        q = qd.as<float>();
        rs = (ptrdiff_t)d;
        cs = 1;
    }
    // q . (y R^T) = (q R) . y: rotate the query like an encode input (pq.rs:276), or just pack it
    RB_TRY(qrot.alloc(nq * d * sizeof(float), st));
    if (pq->proj_dev)
        RB_TRY(launch_project(q, nq, d, rs, cs, pq->proj_dev, 0, qrot.as<float>(), st));
    else
        RB_TRY(launch_pack_rows(q, nq, d, rs, cs, qrot.as<float>(), st));
    RB_TRY(lut.alloc(qstore_lut_floats(M, k, nq) * sizeof(float), st));
    float *dst = out;
    ptrdiff_t dld = out_row_stride;
    if (mem_kind == RB_MEM_HOST) {
        RB_TRY(scores.alloc(nq * n * sizeof(float), st));
        dst = scores.as<float>();
        dld = (ptrdiff_t)n;
    }
    RB_TRY(launch_qstore_dot(s->codes, n, pq->q_dev, M, k, pq->dsub, qrot.as<float>(), nq, s->norms, lut.as<float>(), dst, dld, st));
    if (mem_kind == RB_MEM_HOST) {
        RB_CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)out_row_stride * sizeof(float), dst, n * sizeof(float), n * sizeof(float), nq,
                                      cudaMemcpyDeviceToHost, st));
        RB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return RB_OK;
}

}  // extern "C"

