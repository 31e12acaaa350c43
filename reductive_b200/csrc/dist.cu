// dist.cu — data-parallel Pq k-means across GPUs behind the C ABI; NCCL is called from here (include/reductive_b200.h,
// "multi-GPU" section).
//
// Replaces, for rows sharded over G GPUs, the loop of kmeans_with_centroids (src/kmeans.rs:263-288) around
// kmeans_iteration (src/kmeans.rs:308-327) for all M subquantizers of Pq::train_pq_using (src/pq/pq.rs:201-249).
//
// The two phases of an iteration are sharded on DIFFERENT axes:
//   assignment (kmeans.rs:319)      rows: rank r assigns its n_r rows against all M codebooks (no exchange);
//   update_centroids (kmeans.rs:320) subquantizers: rank r owns M/G subquantizers and adds ALL n rows of each
//     cluster sequentially in row order — exactly the reference's f32 chain (kmeans.rs:185-189), so trained centroids
//     are bit-identical to a one-GPU run (and to the oracle) for any G.  This is the reference's own parallel axis
//     (Rayon over subquantizers, pq.rs:226-241).
// The training rows never change, so each rank receives the column slice x[:, its subquantizers] of every other
// rank ONCE (an all-to-all of the training matrix when the state is created) and keeps both layouts.  Per iteration
// only the assignments travel (n * M bytes in total) and the new centroids are all-gathered (M * k * dsub floats).
// Everything is stream-ordered on the caller's stream.
//
// How the assignments travel: the [M][pitch] code matrix of ALL rows is one virtual address range on every rank
// whose column block of subquantizer m is physical memory of the rank that owns m (CUDA virtual memory management:
// cuMemCreate on the owner, exported as a file descriptor, passed over a Unix socket, cuMemMap on every rank).  The
// assignment kernels of a rank therefore store each code straight into its owner's HBM over NVLink while they run --
// there is no separate exchange step, no receive buffer and no re-assembly; one tiny all-gather orders "all stores
// done" before the owners' updates.  When the window cannot be set up (no peer access, descriptors not passable,
// RB_DIST_P2P=0) the ranks agree to fall back to a grouped ncclSend/Recv of the code blocks.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace rb {
namespace {

// Received assignment blocks -> [m_own][pitch_total] in rank (= row) order.  One launch for all ranks' blocks: block
// (x: 16 KB piece, y: subquantizer, z: source rank).
struct AssembleArgs {
    const unsigned char *src[64];
    unsigned long long src_pitch[64], bytes[64], dst_off[64];
};
__global__ void __launch_bounds__(256) assemble_codes_kernel(const __grid_constant__ AssembleArgs a, unsigned char *__restrict__ dst,
                                                             unsigned long long dst_pitch)
{
    const int r = blockIdx.z, m = blockIdx.y;
    const unsigned long long n = a.bytes[r];
    const unsigned char *s = a.src[r] + (unsigned long long)m * a.src_pitch[r];
    unsigned char *d = dst + (unsigned long long)m * dst_pitch + a.dst_off[r];
    for (unsigned long long i = (unsigned long long)blockIdx.x * 16384 + threadIdx.x * 16ull; i < n && i < (unsigned long long)(blockIdx.x + 1) * 16384;
         i += 256 * 16ull) {
        if (i + 16 <= n && ((reinterpret_cast<unsigned long long>(d + i) & 15) == 0)) {
            *reinterpret_cast<uint4 *>(d + i) = *reinterpret_cast<const uint4 *>(s + i);  // source blocks are 16-byte aligned
        } else {
            for (unsigned long long j = i; j < n && j < i + 16; j++) d[j] = s[j];
        }
    }
}

// NCCL is resolved at run time (dlopen), only when a communicator over more than one device is asked for.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

const NcclApi &nccl()
{
    static NcclApi api = []() {
        NcclApi a;
        // Order: the copy named by RB_NCCL_LIB, the copy the process already carries, the system one.  A host that
        // will load another NCCL later under the same soname (torch's bundled libnccl.so.2) must point RB_NCCL_LIB
        // at that file: the dynamic loader keeps whichever libnccl.so.2 came first for everyone.
        void *h = nullptr;
        if (const char *path = getenv("RB_NCCL_LIB"))
            if (*path) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return a;
#define RB_SYM(name) a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "nccl" #name))
        RB_SYM(GetUniqueId); RB_SYM(CommInitRank); RB_SYM(CommInitAll); RB_SYM(CommDestroy); RB_SYM(GroupStart);
        RB_SYM(GroupEnd); RB_SYM(Send); RB_SYM(Recv); RB_SYM(Broadcast); RB_SYM(AllGather); RB_SYM(GetErrorString);
#undef RB_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send &&
               a.Recv && a.Broadcast && a.AllGather && a.GetErrorString;
        return a;
    }();
    return api;
}

#define RB_NCCL_TRY(expr)                                                                                        \
    do {                                                                                                         \
        ncclResult_t _r = (expr);                                                                                \
        if (_r != ncclSuccess) {                                                                                 \
            ::rb::set_error("%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(_r), __FILE__, __LINE__);      \
            return RB_ERR_NCCL;                                                                                  \
        }                                                                                                        \
    } while (0)


// CUDA driver entry points for virtual memory management, resolved through the runtime (no link-time libcuda).
struct DriverApi {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};

const DriverApi &driver()
{
    static DriverApi api = []() {
        DriverApi a;
        bool all = true;
        auto get = [&](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
                (void)cudaGetLastError();
                all = false;
            }
        };
        get("cuMemCreate", (void **)&a.MemCreate);
        get("cuMemRelease", (void **)&a.MemRelease);
        get("cuMemAddressReserve", (void **)&a.MemAddressReserve);
        get("cuMemAddressFree", (void **)&a.MemAddressFree);
        get("cuMemMap", (void **)&a.MemMap);
        get("cuMemUnmap", (void **)&a.MemUnmap);
        get("cuMemSetAccess", (void **)&a.MemSetAccess);
        get("cuMemExportToShareableHandle", (void **)&a.MemExport);
        get("cuMemImportFromShareableHandle", (void **)&a.MemImport);
        get("cuMemGetAllocationGranularity", (void **)&a.MemGranularity);
        a.ok = all;
        return a;
    }();
    return api;
}

constexpr size_t kWindowGranule = (size_t)2 << 20;  // every column of the window starts on a mapping granule

// The code matrix of all rows, [M][pitch] bytes, as one address range; column block [m_lo[r], m_lo[r + 1]) is memory
// of rank r.
struct CodeWindow {
    CUdeviceptr va = 0;
    size_t va_bytes = 0, pitch = 0;  // pitch: bytes per subquantizer column (a multiple of kWindowGranule)
    std::vector<CUmemGenericAllocationHandle> handle;  // per rank; 0: none
    std::vector<char> mapped;
    bool active = false;
};

void window_release(CodeWindow &w, const std::vector<size_t> &m_lo)
{
    const DriverApi &d = driver();
    if (!d.ok) return;
    for (size_t r = 0; r < w.handle.size(); r++) {
        if (w.mapped.size() > r && w.mapped[r]) d.MemUnmap(w.va + m_lo[r] * w.pitch, (m_lo[r + 1] - m_lo[r]) * w.pitch);
        if (w.handle[r]) d.MemRelease(w.handle[r]);
    }
    if (w.va) d.MemAddressFree(w.va, w.va_bytes);
    w = CodeWindow();
}

// abstract-namespace datagram sockets carry the exported descriptors between the ranks of one node
socklen_t window_addr(sockaddr_un *a, unsigned long long token, int rank)
{
    memset(a, 0, sizeof(*a));
    a->sun_family = AF_UNIX;
    const int len = snprintf(a->sun_path + 1, sizeof(a->sun_path) - 1, "rbw-%016llx-%d", token, rank);
    return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + (size_t)len);
}

bool send_fd(int sock, unsigned long long token, int to_rank, int my_rank, int fd)
{
    sockaddr_un to;
    const socklen_t to_len = window_addr(&to, token, to_rank);
    int payload = my_rank;
    iovec iov{&payload, sizeof(payload)};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    msghdr msg{};
    msg.msg_name = &to;
    msg.msg_namelen = to_len;
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    cmsghdr *cm = CMSG_FIRSTHDR(&msg);
    cm->cmsg_level = SOL_SOCKET;
    cm->cmsg_type = SCM_RIGHTS;
    cm->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(cm), &fd, sizeof(int));
    return sendmsg(sock, &msg, 0) == (ssize_t)sizeof(payload);
}

bool recv_fd(int sock, int *from_rank, int *fd)
{
    int payload = -1;
    iovec iov{&payload, sizeof(payload)};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    msghdr msg{};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    if (recvmsg(sock, &msg, 0) != (ssize_t)sizeof(payload)) return false;
    cmsghdr *cm = CMSG_FIRSTHDR(&msg);
    if (!cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) return false;
    memcpy(fd, CMSG_DATA(cm), sizeof(int));
    *from_rank = payload;
    return true;
}

rb_status require_nccl()
{
    if (!nccl().ok) {
        set_error("libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
        return RB_ERR_NCCL;
    }
    return RB_OK;
}

}  // namespace
}  // namespace rb

using namespace rb;

struct rb_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

// Per-run state of the sharded k-means: both layouts of the training rows and the exchange buffers.
struct rb_kmeans_dist {
    rb_comm *c = nullptr;
    size_t M = 0, k = 0, dsub = 0, d = 0;
    const float *x_local = nullptr;  // borrowed: [n_local, d], row pitch ldx
    ptrdiff_t ldx = 0;
    size_t n_local = 0, n_total = 0;
    std::vector<size_t> n_of, row_off;  // per rank
    std::vector<size_t> m_lo;           // subquantizer ranges: rank r owns [m_lo[r], m_lo[r + 1])
    float *xcol = nullptr;              // owned: [n_total, dcols] with dcols = (m_lo[rank + 1] - m_lo[rank]) * dsub
    unsigned char *codes_local = nullptr, *codes_recv = nullptr, *codes_own = nullptr;
    float *packed_own = nullptr, *loss_all = nullptr;
    double *sumsq_own = nullptr;  // sum ||x_m||^2 of the owned subquantizers over all rows (FP64, computed once)
    float *slabs = nullptr;       // streaming update: the owned columns as subquantizer-major slabs (then xcol is dropped)
    CodeWindow win;               // peer-mapped code matrix (active: assignments are stored straight into their owners' memory)
    unsigned long long *sync_dev = nullptr;  // [W + 1]: scratch of the tiny all-gathers
    int code_width = 1;
    size_t pitch_local = 0, pitch_total = 0;
    size_t m_own() const { return m_lo[c->rank + 1] - m_lo[c->rank]; }
};

namespace {

// every rank contributes one 64-bit value and learns all of them (doubles as a barrier of the ranks' streams)
rb_status exchange_u64(rb_kmeans_dist *h, cudaStream_t st, unsigned long long mine, std::vector<unsigned long long> &all)
{
    const int W = h->c->world;
    RB_CUDA_TRY(cudaMemcpyAsync(h->sync_dev + W, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    RB_NCCL_TRY(nccl().AllGather(h->sync_dev + W, h->sync_dev, 1, ncclUint64, h->c->comm, st));
    all.assign(W, 0);
    RB_CUDA_TRY(cudaMemcpyAsync(all.data(), h->sync_dev, (size_t)W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    RB_CUDA_TRY(cudaStreamSynchronize(st));
    return RB_OK;
}

// Builds the peer-mapped code matrix (see the file header).  Collective.  Any local failure is voted on, so that either
// every rank ends with win.active or none does; only a failing collective is an error.
rb_status window_setup(rb_kmeans_dist *h, cudaStream_t st)
{
    const int W = h->c->world, me = h->c->rank;
    const size_t M = h->M, cw = (size_t)h->code_width, m_own = h->m_own();
    CodeWindow &w = h->win;
    const DriverApi &d = driver();
    const char *env = getenv("RB_DIST_P2P");
    bool ok = d.ok && !(env && env[0] == '0');
    w.pitch = ceil_div(h->pitch_total * cw, kWindowGranule) * kWindowGranule;
    w.handle.assign(W, 0);
    w.mapped.assign(W, 0);
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = h->c->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    struct Fd {  // closed on every way out, also when a collective fails
        int v = -1;
        ~Fd()
        {
            if (v >= 0) close(v);
        }
        Fd &operator=(int x)
        {
            v = x;
            return *this;
        }
        operator int() const { return v; }
        int *ptr() { return &v; }
    } my_fd, sock;
    if (ok) {
        size_t g = 0;
        ok = d.MemGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) == CUDA_SUCCESS && g != 0 && kWindowGranule % g == 0;
    }
    if (ok && m_own) {
        ok = d.MemCreate(&w.handle[me], m_own * w.pitch, &prop, 0) == CUDA_SUCCESS;
        if (!ok) w.handle[me] = 0;
        if (ok) ok = d.MemExport(my_fd.ptr(), w.handle[me], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
    }
    // a socket name nobody else uses: rank 0's random token
    unsigned long long token = 0;
    if (me == 0) {
        FILE *f = fopen("/dev/urandom", "rb");
        if (!f || fread(&token, sizeof(token), 1, f) != 1) token = ((unsigned long long)getpid() << 32) ^ (unsigned long long)clock();
        if (f) fclose(f);
    }
    std::vector<unsigned long long> all;
    RB_TRY(exchange_u64(h, st, token, all));
    token = all[0];
    if (ok) {
        sock = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
        sockaddr_un a;
        const socklen_t alen = window_addr(&a, token, me);
        timeval tv{20, 0};
        ok = sock >= 0 && bind(sock, reinterpret_cast<sockaddr *>(&a), alen) == 0 &&
             setsockopt(sock, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv)) == 0;
    }
    RB_TRY(exchange_u64(h, st, ok ? 1ull : 0ull, all));  // everyone is bound (or someone gave up)
    bool go = true;
    for (int r = 0; r < W; r++) go = go && all[r] != 0;
    if (go) {
        if (m_own)
            for (int r = 0; r < W && ok; r++)
                if (r != me) ok = send_fd(sock, token, r, me, my_fd);
        for (int r = 0; r < W && ok; r++) {
            if (r == me || h->m_lo[r + 1] == h->m_lo[r]) continue;  // one descriptor from every other owner, any order
            int from = -1, fd = -1;
            if (!recv_fd(sock, &from, &fd)) {
                ok = false;
                break;
            }
            if (from < 0 || from >= W || from == me || w.handle[from] != 0) {
                close(fd);
                ok = false;
                break;
            }
            ok = d.MemImport(&w.handle[from], (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
            if (!ok) w.handle[from] = 0;
            close(fd);
        }
        if (ok) {
            w.va_bytes = M * w.pitch;
            ok = d.MemAddressReserve(&w.va, w.va_bytes, kWindowGranule, 0, 0) == CUDA_SUCCESS;
            if (!ok) w.va = 0;
        }
        for (int r = 0; r < W && ok; r++) {
            const size_t mr = h->m_lo[r + 1] - h->m_lo[r];
            if (!mr) continue;
            ok = d.MemMap(w.va + h->m_lo[r] * w.pitch, mr * w.pitch, 0, w.handle[r], 0) == CUDA_SUCCESS;
            if (ok) w.mapped[r] = 1;
        }
        if (ok) {
            CUmemAccessDesc acc;
            memset(&acc, 0, sizeof(acc));
            acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            acc.location.id = h->c->device;
            acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            ok = d.MemSetAccess(w.va, w.va_bytes, &acc, 1) == CUDA_SUCCESS;  // peers' memory: needs P2P between the devices
        }
        if (ok && m_own)
            ok = cudaMemsetAsync(reinterpret_cast<void *>(w.va + h->m_lo[me] * w.pitch), 0, m_own * w.pitch, st) == cudaSuccess;
    }
    RB_TRY(exchange_u64(h, st, (go && ok) ? 1ull : 0ull, all));
    bool active = true;
    for (int r = 0; r < W; r++) active = active && all[r] != 0;
    if (!active) {
        (void)cudaGetLastError();
        window_release(w, h->m_lo);
    } else {
        w.active = true;
    }
    return RB_OK;
}

}  // namespace

extern "C" {

rb_status rb_dist_subquantizer_range(size_t n_subquantizers, int rank, int world, size_t *m_begin, size_t *m_end)
{
    if (world <= 0 || rank < 0 || rank >= world || !m_begin || !m_end) {
        set_error("bad rank / world (%d / %d)", rank, world);
        return RB_ERR_INVALID;
    }
    *m_begin = n_subquantizers * (size_t)rank / (size_t)world;
    *m_end = n_subquantizers * (size_t)(rank + 1) / (size_t)world;
    return RB_OK;
}

rb_status rb_comm_unique_id(void *id_out, size_t len)
{
    if (!id_out || len < sizeof(ncclUniqueId)) {
        set_error("id buffer must hold %zu bytes", sizeof(ncclUniqueId));
        return RB_ERR_INVALID;
    }
    RB_TRY(require_nccl());
    ncclUniqueId id;
    RB_NCCL_TRY(nccl().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return RB_OK;
}

rb_status rb_comm_create(const void *id, int rank, int world, rb_comm **out)
{
    if (!out) return RB_ERR_INVALID;
    *out = nullptr;
    if ((!id && world != 1) || world <= 0 || rank < 0 || rank >= world) {
        set_error("bad communicator arguments (rank %d of %d)", rank, world);
        return RB_ERR_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device visible; reductive_b200 has no CPU fallback");
        return RB_ERR_NO_DEVICE;
    }
    rb_comm *c = new rb_comm();
    c->rank = rank;
    c->world = world;
    cudaGetDevice(&c->device);
    if (world == 1) {  // a single rank exchanges nothing: no NCCL communicator behind it
        *out = c;
        return RB_OK;
    }
    if (require_nccl() != RB_OK) {
        delete c;
        return RB_ERR_NCCL;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const ncclResult_t r = nccl().CommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
        delete c;
        return RB_ERR_NCCL;
    }
    *out = c;
    return RB_OK;
}

void rb_comm_destroy(rb_comm *c)
{
    if (!c) return;
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

int rb_comm_rank(const rb_comm *c) { return c ? c->rank : -1; }
int rb_comm_world(const rb_comm *c) { return c ? c->world : 0; }

void rb_kmeans_dist_destroy(rb_kmeans_dist *h)
{
    if (!h) return;
    cudaFree(h->xcol);
    cudaFree(h->codes_local);
    cudaFree(h->codes_recv);
    cudaFree(h->codes_own);
    cudaFree(h->packed_own);
    cudaFree(h->loss_all);
    cudaFree(h->sumsq_own);
    cudaFree(h->slabs);
    cudaFree(h->sync_dev);
    if (h->win.va || !h->win.handle.empty()) {
        cudaDeviceSynchronize();  // nothing may still be storing into the window
        window_release(h->win, h->m_lo);
    }
    delete h;
}

rb_status rb_kmeans_dist_create(rb_comm *c, const float *x_local, size_t n_local, ptrdiff_t x_row_stride, size_t M,
                                size_t k, size_t dsub, void *stream, rb_kmeans_dist **out)
{
    if (!out) return RB_ERR_INVALID;
    *out = nullptr;
    if (!c || (n_local && !x_local)) {
        set_error("NULL argument");
        return RB_ERR_INVALID;
    }
    if (M == 0 || k == 0 || dsub == 0 || x_row_stride < (ptrdiff_t)(M * dsub)) {
        set_error("bad shape (M=%zu k=%zu dsub=%zu row stride %td)", M, k, dsub, x_row_stride);
        return RB_ERR_SHAPE;
    }
    RB_TRY(require_nccl());
    cudaStream_t st = (cudaStream_t)stream;
    const int W = c->world, me = c->rank;
    rb_kmeans_dist *h = new rb_kmeans_dist();
    h->c = c;
    h->M = M; h->k = k; h->dsub = dsub; h->d = M * dsub;
    h->x_local = x_local;
    h->ldx = x_row_stride;
    h->n_local = n_local;
    h->code_width = rb_kmeans_code_width(k);
    h->m_lo.resize(W + 1);
    for (int r = 0; r <= W; r++) h->m_lo[r] = M * (size_t)r / (size_t)W;
    auto body = [&]() -> rb_status {
        // row counts of every rank (rank r's rows follow rank r - 1's in the reference's row order)
        std::vector<unsigned long long> cnt(W);
        if (W == 1) {
            cnt[0] = n_local;
        } else {
            RB_TRY(require_nccl());
            unsigned long long *cnt_dev = nullptr;
            RB_CUDA_TRY(cudaMalloc(&cnt_dev, (size_t)(W + 1) * sizeof(unsigned long long)));
            const unsigned long long mine = n_local;
            rb_status s = [&]() -> rb_status {
                RB_CUDA_TRY(cudaMemcpyAsync(cnt_dev + W, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
                RB_NCCL_TRY(nccl().AllGather(cnt_dev + W, cnt_dev, 1, ncclUint64, c->comm, st));
                RB_CUDA_TRY(cudaMemcpyAsync(cnt.data(), cnt_dev, (size_t)W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                RB_CUDA_TRY(cudaStreamSynchronize(st));
                return RB_OK;
            }();
            cudaFree(cnt_dev);
            RB_TRY(s);
        }
        h->n_of.resize(W);
        h->row_off.resize(W + 1);
        h->row_off[0] = 0;
        for (int r = 0; r < W; r++) {
            h->n_of[r] = (size_t)cnt[r];
            h->row_off[r + 1] = h->row_off[r] + h->n_of[r];
        }
        h->n_total = h->row_off[W];
        h->pitch_local = rb_kmeans_code_pitch(n_local);
        h->pitch_total = rb_kmeans_code_pitch(h->n_total);
        const size_t m_own = h->m_own(), dcols = m_own * dsub, cw = (size_t)h->code_width;
        size_t recv_bytes = 0;
        for (int r = 0; r < W; r++) recv_bytes += m_own * rb_kmeans_code_pitch(h->n_of[r]) * cw;
        const bool stream_upd = kmeans_stream_enabled() && cw == 1 && stream_update_supported(k, dsub);
        if (W == 1) RB_CUDA_TRY(cudaMalloc(&h->codes_local, M * h->pitch_local * cw + 16));
        RB_CUDA_TRY(cudaMalloc(&h->packed_own, (rb_kmeans_packed_len(m_own ? m_own : 1, k, dsub)) * sizeof(float)));
        RB_CUDA_TRY(cudaMalloc(&h->loss_all, M * sizeof(float)));
        RB_CUDA_TRY(cudaMalloc(&h->sumsq_own, (m_own ? m_own : 1) * sizeof(double)));
        if (W == 1) {
            // one rank: nothing to exchange; the rows stay where they are (row-major), slabs are built from them
            RB_TRY(launch_sumsq64(x_local, n_local, x_row_stride, M, dsub, h->sumsq_own, st));
            if (stream_upd) {
                RB_CUDA_TRY(cudaMalloc(&h->slabs, slab_floats(n_local, h->d) * sizeof(float)));
                RB_TRY(launch_build_slabs(x_local, n_local, x_row_stride, M, dsub, h->slabs, st));
            }
            return RB_OK;
        }
        RB_CUDA_TRY(cudaMalloc(&h->xcol, (h->n_total * dcols + 4) * sizeof(float)));
        RB_CUDA_TRY(cudaMalloc(&h->sync_dev, (size_t)(W + 1) * sizeof(unsigned long long)));
        RB_TRY(window_setup(h, st));
        if (!h->win.active) {  // exchange by ncclSend/Recv: local assignments, receive blocks, assembled own columns
            RB_CUDA_TRY(cudaMalloc(&h->codes_local, M * h->pitch_local * cw + 16));
            RB_CUDA_TRY(cudaMalloc(&h->codes_recv, recv_bytes + 16));
            RB_CUDA_TRY(cudaMalloc(&h->codes_own, m_own * h->pitch_total * cw + 16));
            RB_CUDA_TRY(cudaMemsetAsync(h->codes_own, 0, m_own * h->pitch_total * cw + 16, st));
        }
        // one-time all-to-all of the training matrix: rank s receives x[rows of r, columns of s] from every r
        float *sendbuf = nullptr;
        RB_CUDA_TRY(pool_malloc((void **)&sendbuf, (n_local * h->d + 4) * sizeof(float), st));
        rb_status s2 = [&]() -> rb_status {
            std::vector<float *> send_at(W);
            size_t off = 0;
            for (int r = 0; r < W; r++) {
                const size_t cols = (h->m_lo[r + 1] - h->m_lo[r]) * dsub;
                send_at[r] = sendbuf + off;
                if (cols && n_local)
                    RB_CUDA_TRY(cudaMemcpy2DAsync(send_at[r], cols * sizeof(float), x_local + h->m_lo[r] * dsub,
                                                  (size_t)x_row_stride * sizeof(float), cols * sizeof(float), n_local,
                                                  cudaMemcpyDeviceToDevice, st));
                off += n_local * cols;
            }
            RB_NCCL_TRY(nccl().GroupStart());
            for (int r = 0; r < W; r++) {
                const size_t cols_r = (h->m_lo[r + 1] - h->m_lo[r]) * dsub;
                if (n_local * cols_r) RB_NCCL_TRY(nccl().Send(send_at[r], n_local * cols_r, ncclFloat, r, c->comm, st));
                if (h->n_of[r] * dcols)
                    RB_NCCL_TRY(nccl().Recv(h->xcol + h->row_off[r] * dcols, h->n_of[r] * dcols, ncclFloat, r, c->comm, st));
            }
            RB_NCCL_TRY(nccl().GroupEnd());
            return RB_OK;
        }();
        cudaFreeAsync(sendbuf, st);
        RB_TRY(s2);
        RB_TRY(launch_sumsq64(h->xcol, h->n_total, (ptrdiff_t)dcols, m_own, dsub, h->sumsq_own, st));
        if (stream_upd && m_own) {  // the update streams slabs: build them once, the column shard is not needed again
            RB_CUDA_TRY(cudaMalloc(&h->slabs, slab_floats(h->n_total, dcols) * sizeof(float)));
            RB_TRY(launch_build_slabs(h->xcol, h->n_total, (ptrdiff_t)dcols, m_own, dsub, h->slabs, st));
            RB_CUDA_TRY(cudaStreamSynchronize(st));
            cudaFree(h->xcol);
            h->xcol = nullptr;
        }
        return RB_OK;
    };
    (void)me;
    const rb_status s = body();
    if (s != RB_OK) {
        rb_kmeans_dist_destroy(h);
        return s;
    }
    *out = h;
    return RB_OK;
}

rb_status rb_kmeans_dist_iterate(rb_kmeans_dist *h, float *centroids, float *loss_or_null, void *stream)
{
    if (!h || !centroids) {
        set_error("NULL argument");
        return RB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rb_comm *c = h->c;
    const int W = c->world, me = c->rank;
    const size_t M = h->M, k = h->k, dsub = h->dsub, cw = (size_t)h->code_width;
    const size_t m_own = h->m_own(), dcols = m_own * dsub;
    // 1. cluster_assignments of the local rows against all M codebooks (kmeans.rs:319)
    const bool window = W > 1 && h->win.active;
    if (window) {
        // ... stored straight into the owners' memory: this rank's rows of every column of the peer-mapped matrix
        unsigned char *base = reinterpret_cast<unsigned char *>(h->win.va);
        RB_TRY(kmeans_assign_strided(h->x_local, h->n_local, h->ldx, centroids, M, k, dsub, base + h->row_off[me] * cw,
                                     (ptrdiff_t)(h->win.pitch / cw), st));
        // every rank's stores are complete once all ranks have passed this point of their streams
        RB_NCCL_TRY(nccl().AllGather(h->sync_dev + W, h->sync_dev, 1, ncclUint64, c->comm, st));
    } else {
        RB_TRY(rb_kmeans_assign(h->x_local, h->n_local, h->ldx, centroids, M, k, dsub, h->codes_local, stream));
    }
    // 2. the assignments of subquantizers [m_lo[s], m_lo[s+1]) go to rank s (column-major blocks are contiguous)
    const unsigned char *codes_own = h->codes_own;
    size_t pitch_own = h->pitch_total;
    const float *x_own = h->xcol;
    ptrdiff_t ld_own = (ptrdiff_t)dcols;
    if (W == 1) {  // one rank owns everything: the local assignments ARE the owned ones
        codes_own = h->codes_local;
        pitch_own = h->pitch_local;
        x_own = h->x_local;
        ld_own = h->ldx;
    } else if (window) {  // the owned columns of the window are local memory and already complete
        codes_own = reinterpret_cast<const unsigned char *>(h->win.va) + h->m_lo[me] * h->win.pitch;
        pitch_own = h->win.pitch / cw;
    } else {
    std::vector<unsigned char *> recv_at(W);
    {
        size_t off = 0;
        for (int r = 0; r < W; r++) {
            recv_at[r] = h->codes_recv + off;
            off += m_own * rb_kmeans_code_pitch(h->n_of[r]) * cw;
        }
    }
    RB_NCCL_TRY(nccl().GroupStart());
    for (int r = 0; r < W; r++) {
        const size_t send_bytes = (h->m_lo[r + 1] - h->m_lo[r]) * h->pitch_local * cw;
        const size_t recv_bytes = m_own * rb_kmeans_code_pitch(h->n_of[r]) * cw;
        if (send_bytes && h->n_local)
            RB_NCCL_TRY(nccl().Send(h->codes_local + h->m_lo[r] * h->pitch_local * cw, send_bytes, ncclUint8, r, c->comm, st));
        if (recv_bytes && h->n_of[r]) RB_NCCL_TRY(nccl().Recv(recv_at[r], recv_bytes, ncclUint8, r, c->comm, st));
    }
    RB_NCCL_TRY(nccl().GroupEnd());
    // rows of all ranks in rank order = the reference's row order: [m_own][pitch_total]
    if (m_own && W <= 64) {
        AssembleArgs a;
        unsigned long long most = 0;
        for (int r = 0; r < 64; r++) {
            a.src[r] = r < W ? recv_at[r] : nullptr;
            a.src_pitch[r] = r < W ? rb_kmeans_code_pitch(h->n_of[r]) * cw : 0;
            a.bytes[r] = r < W ? h->n_of[r] * cw : 0;
            a.dst_off[r] = r < W ? h->row_off[r] * cw : 0;
            most = a.bytes[r] > most ? a.bytes[r] : most;
        }
        if (most) {
            assemble_codes_kernel<<<dim3((unsigned)ceil_div((size_t)most, (size_t)16384), (unsigned)m_own, (unsigned)W), 256, 0, st>>>(
                a, h->codes_own, h->pitch_total * cw);
            RB_LAUNCH_CHECK();
        }
    } else {
        for (int r = 0; r < W; r++)
            if (m_own && h->n_of[r])
                RB_CUDA_TRY(cudaMemcpy2DAsync(h->codes_own + h->row_off[r] * cw, h->pitch_total * cw, recv_at[r],
                                              rb_kmeans_code_pitch(h->n_of[r]) * cw, h->n_of[r] * cw, m_own,
                                              cudaMemcpyDeviceToDevice, st));
    }
    }
    // 3. update_centroids of the owned subquantizers over ALL rows, in row order (kmeans.rs:166-198)
    if (m_own) {
        if (h->slabs)
            RB_TRY(launch_ordered_stream(h->slabs, h->n_total, codes_own, pitch_own, m_own, k, dsub, h->packed_own, st));
        else
            RB_TRY(launch_kmeans_accumulate(x_own, h->n_total, ld_own, cw == 1 ? codes_own : nullptr,
                                            cw == 4 ? reinterpret_cast<const uint32_t *>(codes_own) : nullptr, pitch_own, m_own, k,
                                            dsub, nullptr, h->packed_own, kmeans_update_mode() != 0 ? 1 : 0, st));
        RB_TRY(launch_kmeans_finalize(h->packed_own, m_own, k, dsub, h->n_total, centroids + h->m_lo[me] * k * dsub,
                                      h->loss_all + h->m_lo[me], st, h->sumsq_own));
    }
    // 4. everyone gets everyone's new centroids (and, when asked for, losses): one in-place all-gather when the
    //    subquantizers divide evenly, else one broadcast per rank
    const bool even = M % (size_t)W == 0;
    if (W == 1) {
        // nothing to exchange
    } else if (even) {
        RB_NCCL_TRY(nccl().AllGather(centroids + h->m_lo[me] * k * dsub, centroids, m_own * k * dsub, ncclFloat, c->comm, st));
        if (loss_or_null) RB_NCCL_TRY(nccl().AllGather(h->loss_all + h->m_lo[me], h->loss_all, m_own, ncclFloat, c->comm, st));
    } else {
        RB_NCCL_TRY(nccl().GroupStart());
        for (int r = 0; r < W; r++) {
            const size_t mr = h->m_lo[r + 1] - h->m_lo[r];
            if (!mr) continue;
            float *cen_r = centroids + h->m_lo[r] * k * dsub;
            RB_NCCL_TRY(nccl().Broadcast(cen_r, cen_r, mr * k * dsub, ncclFloat, r, c->comm, st));
            if (loss_or_null)
                RB_NCCL_TRY(nccl().Broadcast(h->loss_all + h->m_lo[r], h->loss_all + h->m_lo[r], mr, ncclFloat, r, c->comm, st));
        }
        RB_NCCL_TRY(nccl().GroupEnd());
    }
    if (loss_or_null)
        RB_CUDA_TRY(cudaMemcpyAsync(loss_or_null, h->loss_all, M * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return RB_OK;
}

int rb_kmeans_dist_peer_window(const rb_kmeans_dist *h) { return h && h->win.active ? 1 : 0; }

rb_status rb_pq_train_dist(rb_comm *c, const float *instances_local, size_t n_local, size_t n_total, size_t d,
                           ptrdiff_t row_stride, size_t n_subquantizers, uint32_t n_subquantizer_bits, size_t n_iterations,
                           size_t n_attempts, const float *initial_centroids, float *loss_out, int mem_kind, void *stream,
                           rb_pq **out)
{
    if (!out) return RB_ERR_INVALID;
    *out = nullptr;
    const size_t M = n_subquantizers;
    RB_TRY(rb_check_quantizer_invariants(M, n_subquantizer_bits, n_iterations, n_attempts, n_total, d, nullptr));  // pq.rs:213-219
    if (!c || !initial_centroids || (n_local && !instances_local)) {
        set_error("NULL argument");
        return RB_ERR_INVALID;
    }
    const size_t k = (size_t)1 << n_subquantizer_bits;
    if (k >= n_total) {  // kmeans.rs:62-67
        set_error("Cannot pick more centroids than instances: %zu instances, %zu centroids", n_total, k);
        return RB_ERR_K_MEANS_K;
    }
    if (row_stride < (ptrdiff_t)d) {
        set_error("row stride %td < %zu columns", row_stride, d);
        return RB_ERR_SHAPE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t dsub = d / M, qn = M * k * dsub;
    float *xdev = nullptr;
    const float *x = instances_local;
    ptrdiff_t ldx = row_stride;
    if (mem_kind == RB_MEM_HOST) {
        RB_CUDA_TRY(cudaMalloc(&xdev, (n_local * d + 4) * sizeof(float)));
        if (n_local)
            RB_CUDA_TRY(cudaMemcpy2DAsync(xdev, d * sizeof(float), instances_local, (size_t)row_stride * sizeof(float),
                                          d * sizeof(float), n_local, cudaMemcpyHostToDevice, st));
        x = xdev;
        ldx = (ptrdiff_t)d;
    }
    rb_kmeans_dist *h = nullptr;
    float *cen = nullptr, *loss_dev = nullptr;
    std::vector<float> best_q(qn), cand_q(qn), best_loss(M, 0.f), cand_loss(M, 0.f);
    auto body = [&]() -> rb_status {
        RB_TRY(rb_kmeans_dist_create(c, x, n_local, ldx, M, k, dsub, stream, &h));
        if (h->n_total != n_total) {
            set_error("the ranks hold %zu rows in total, n_total says %zu", h->n_total, n_total);
            return RB_ERR_SHAPE;
        }
        RB_CUDA_TRY(cudaMalloc(&cen, qn * sizeof(float)));
        RB_CUDA_TRY(cudaMalloc(&loss_dev, M * sizeof(float)));
        for (size_t a = 0; a < n_attempts; a++) {  // pq.rs:168-183, all subquantizers advance together
            RB_CUDA_TRY(cudaMemcpyAsync(cen, initial_centroids + a * qn, qn * sizeof(float), cudaMemcpyHostToDevice, st));
            for (size_t it = 0; it < n_iterations; it++)  // kmeans.rs:279-284
                RB_TRY(rb_kmeans_dist_iterate(h, cen, it + 1 == n_iterations ? loss_dev : nullptr, stream));
            RB_CUDA_TRY(cudaMemcpyAsync(cand_q.data(), cen, qn * sizeof(float), cudaMemcpyDeviceToHost, st));
            RB_CUDA_TRY(cudaMemcpyAsync(cand_loss.data(), loss_dev, M * sizeof(float), cudaMemcpyDeviceToHost, st));
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
        return RB_OK;
    };
    const rb_status s = body();
    rb_kmeans_dist_destroy(h);
    cudaFree(cen);
    cudaFree(loss_dev);
    cudaFree(xdev);
    RB_TRY(s);
    if (loss_out) memcpy(loss_out, best_loss.data(), M * sizeof(float));
    return rb_pq_create(best_q.data(), M, k, dsub, nullptr, out);
}

// One process, several GPUs: one host thread per device, communicators from ncclCommInitAll, each thread runs the
// rank's part of rb_pq_train_dist on its contiguous block of rows.  instances: HOST [n, d].
rb_status rb_pq_train_multi(const int *devices, int n_devices, const float *instances, size_t n, size_t d,
                            ptrdiff_t row_stride, size_t n_subquantizers, uint32_t n_subquantizer_bits, size_t n_iterations,
                            size_t n_attempts, const float *initial_centroids, float *loss_out, rb_pq **out)
{
    if (!out) return RB_ERR_INVALID;
    *out = nullptr;
    if (!devices || n_devices <= 0 || !instances || !initial_centroids) {
        set_error("NULL argument / no devices");
        return RB_ERR_INVALID;
    }
    RB_TRY(rb_check_quantizer_invariants(n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, n, d, nullptr));
    RB_TRY(require_nccl());
    const int W = n_devices;
    std::vector<ncclComm_t> comms(W);
    RB_NCCL_TRY(nccl().CommInitAll(comms.data(), W, devices));
    std::vector<rb_status> status(W, RB_OK);
    std::vector<std::string> errors(W);
    std::vector<rb_pq *> pqs(W, nullptr);
    std::vector<std::vector<float>> losses(W, std::vector<float>(n_subquantizers));
    const size_t per = (n + (size_t)W - 1) / (size_t)W;
    std::vector<std::thread> threads;
    for (int r = 0; r < W; r++) {
        threads.emplace_back([&, r]() {
            rb_comm c;
            c.comm = comms[r];
            c.rank = r;
            c.world = W;
            c.device = devices[r];
            if (cudaSetDevice(devices[r]) != cudaSuccess) {
                status[r] = RB_ERR_CUDA;
                errors[r] = "cudaSetDevice failed";
                return;
            }
            const size_t lo = per * (size_t)r < n ? per * (size_t)r : n;
            const size_t hi = lo + per < n ? lo + per : n;
            cudaStream_t st = nullptr;
            cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            status[r] = rb_pq_train_dist(&c, instances + (ptrdiff_t)lo * row_stride, hi - lo, n, d, row_stride, n_subquantizers,
                                         n_subquantizer_bits, n_iterations, n_attempts, initial_centroids, losses[r].data(),
                                         RB_MEM_HOST, st, &pqs[r]);
            if (status[r] != RB_OK) errors[r] = rb_last_error_message();
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        });
    }
    for (auto &t : threads) t.join();
    for (int r = 0; r < W; r++) nccl().CommDestroy(comms[r]);
    rb_status s = RB_OK;
    for (int r = 0; r < W; r++)
        if (status[r] != RB_OK && s == RB_OK) {
            s = status[r];
            set_error("device %d: %s", devices[r], errors[r].c_str());
        }
    // every rank holds the same quantizer: keep the first device's handle
    for (int r = 1; r < W; r++) {
        if (pqs[r]) {
            cudaSetDevice(devices[r]);
            rb_pq_destroy(pqs[r]);
        }
    }
    cudaSetDevice(devices[0]);
    if (s != RB_OK) {
        if (pqs[0]) rb_pq_destroy(pqs[0]);
        return s;
    }
    if (loss_out) memcpy(loss_out, losses[0].data(), n_subquantizers * sizeof(float));
    *out = pqs[0];
    return RB_OK;
}

}  // extern "C"
