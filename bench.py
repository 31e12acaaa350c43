#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric: quantize_batch vectors/s (config C2: 2M x 300 f32,
30 subquantizers x 256 centroids, u8 codes) on N B200s, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU algorithm (oracle port)

A "step" is one quantize_batch pass over one device-resident batch.  With N > 1 every rank encodes its own
2M-row shard (rows are independent: no data-path collective) -> "scaling": "weak"; `value` = rows all ranks
encoded / the slowest rank's device time.  `e2e` is the same metric through the C ABI with HOST (pinned) buffers:
host->device copy of the vectors and device->host copy of the codes inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "quantize_batch_vectors_per_s"
UNIT = "vectors/s"
N_ROWS, M, K_CENTROIDS, DSUB = 2_000_000, 30, 256, 10
D = M * DSUB
# dram__bytes_read.sum + dram__bytes_write.sum of encode_tc_kernel<10, 0> on this workload (ncu --set full, round-2 build:
# profiles/r2_encode_tc_final_ncu.txt)
ENCODE_DRAM_BYTES_PER_LAUNCH = 2.4537e9 + 0.0708e9
WORKLOAD = "C2: quantize_batch 2M x 300 f32 N(0,1), 30 subquantizers x 256 centroids, u8 codes"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region is tens of milliseconds, so the
    sampler polls NVML directly (every ~1 ms) from a thread; `nvidia-smi -lms` (B200_PROFILING.md's clocks line)
    cannot start that fast and is only the fallback."""

    def __init__(self, gpu_index: int):
        self.gpu, self.samples, self.reasons, self.mx = gpu_index, [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            self.ready.set()
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.001)
        except Exception as e:  # noqa: BLE001 - reported in the JSON line
            self.err = repr(e)
            self.ready.set()

    def start(self):
        self.ready = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        self.ready.wait(timeout=10)
        self.samples.clear()

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "samples": 0,
                    "reasons": ["clock sampling unavailable: " + str(self.err)]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.mx, "samples": len(self.samples),
                "reasons": sorted(self.reasons)}


def _codebook(seed=1):
    return np.random.default_rng(seed).normal(size=(M, K_CENTROIDS, DSUB)).astype(np.float32)


# --------------------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    """The reference's own CPU algorithm for the path (oracle port: the reference is Rust and cannot be built in
    this image).  Rows are sharded over all host cores; each thread runs the reference's sequential
    quantize_batch loop (primitives.rs:90) on its block — more generous than the reference, whose
    quantize_batch is single-threaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    o = orc.get()
    cores = os.cpu_count() or 1
    q = _codebook()
    sample = int(os.environ.get("RB_REF_SAMPLE_ROWS", 500_000))
    x = np.random.default_rng(2).normal(size=(sample, D)).astype(np.float32)
    for _ in range(args.warmup):
        o.quantize_batch(q, None, x[: max(1024, sample // 8)], np.uint8, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.quantize_batch(q, None, x, np.uint8, n_threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    desc = f"{sample} of the {N_ROWS} rows per step, {cores} threads (row-sharded), AVX2+FMA oracle port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import reductive_b200 as rb
    from reductive_b200 import _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    # stdout carries exactly ONE line, the JSON: libraries that print banners there (NCCL's version line) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def log(msg):  # progress on stderr (RB_BENCH_VERBOSE=1): where a multi-rank run stands
        if os.environ.get("RB_BENCH_VERBOSE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    log("start")
    peaks = _peaks()
    q = _codebook()
    pq = rb.Pq(None, q)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    x = torch.randn((N_ROWS, D), generator=g, device=dev, dtype=torch.float32)  # 2.4 GB > 126 MB L2
    codes = torch.empty((N_ROWS, M), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        pq.quantize_batch_into(x, codes)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = rb.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    launches = rb.kernel_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * N_ROWS / (ms_per_step * 1e-3)

    log("device-resident steps done")
    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region -------
    # The staging buffers are allocated and driven from the CPUs NVML reports as local to this rank's GPU (NUMA
    # placement of the pinned pages matters once several ranks copy at the same time); restored afterwards.
    prev_affinity = _bind_to_gpu_cpus(local_rank)
    e2e_rows = N_ROWS
    xh = torch.empty((e2e_rows, D), dtype=torch.float32, pin_memory=True)
    xh.copy_(x[:e2e_rows])
    ch = torch.empty((e2e_rows, M), dtype=torch.uint8, pin_memory=True)
    xh_np, ch_np = xh.numpy(), ch.numpy()
    e2e_steps = max(1, min(args.steps, 5))
    pq.quantize_batch_into(xh_np, ch_np)  # warm-up (allocations, stream creation)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pq.quantize_batch_into(xh_np, ch_np)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_rows / float(te.item())
    e2e_ok = bool(torch.equal(ch.to(dev), codes[:e2e_rows]))
    # the same call on a plain numpy array (pageable memory, what an ndarray caller hands over): the library stages
    # it through pinned buffers of its own
    xn = np.empty((e2e_rows, D), np.float32)
    xn[:] = xh_np
    cn = np.empty((e2e_rows, M), np.uint8)
    pq.quantize_batch_into(xn, cn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pq.quantize_batch_into(xn, cn)
    pg_s = (time.perf_counter() - t0) / e2e_steps
    tp = torch.tensor([pg_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    e2e_pageable = world * e2e_rows / float(tp.item())
    e2e_pageable_ok = bool(np.array_equal(cn, ch_np))
    # ... and from ONE process over all GPUs of the box (rb_pq_create_multi splits the rows; rank 0 only, others idle)
    one_process = None
    if torch.cuda.device_count() > 1:
        barrier()
        if rank == 0:
            ndev = world if world > 1 else torch.cuda.device_count()
            pqm = rb.Pq(None, q, devices=list(range(ndev)))
            rows_m = 1_000_000 * ndev
            xm = np.empty((rows_m, D), np.float32)
            for i in range(0, rows_m, 1_000_000):
                xm[i:i + 1_000_000] = xn[:1_000_000]
            cm = np.empty((rows_m, M), np.uint8)
            pqm.quantize_batch_into(xm, cm)
            t0 = time.perf_counter()
            for _ in range(2):
                pqm.quantize_batch_into(xm, cm)
            dt = (time.perf_counter() - t0) / 2
            one_process = {"devices": ndev, "rows": rows_m, "value": rows_m / dt, "unit": UNIT, "buffers": "pageable numpy",
                           "codes_match": bool(np.array_equal(cm[:1_000_000], cn[:1_000_000]))}
            del xm, cm, pqm
        barrier()
    del xn, cn
    if prev_affinity:
        os.sched_setaffinity(0, prev_affinity)

    log("e2e done")
    # ---- secondary measurements (same JSON line, "extra"): reconstruct_batch and Pq k-means sec/iter ------
    extra = {}
    rec = torch.empty((N_ROWS, D), dtype=torch.float32, device=dev)
    for _ in range(3):
        pq.reconstruct_batch_into(codes, rec)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    for _ in range(args.steps):
        pq.reconstruct_batch_into(codes, rec)
    r1.record(stream)
    barrier()
    tr = torch.tensor([r0.elapsed_time(r1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    rec_ms = float(tr.item())
    rec_gbs = N_ROWS * (M + 4 * D) / (rec_ms * 1e-3) / 1e9
    extra["reconstruct_batch"] = {
        "workload": "C2 decode: 2M x 30 u8 codes -> 2M x 300 f32", "ms": rec_ms,
        "vectors_per_s": world * N_ROWS / (rec_ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": rec_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": rec_gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_vector": M + 4 * D}}
    del rec
    log("reconstruct done")
    # C3: Pq k-means 1M x 768, 96 x 256 centroids, 25 iterations; rows sharded over the ranks (strong scaling).
    # Sharded mode (csrc/dist.cu): assignment by rows, ordered centroid update by subquantizers, NCCL from the C++
    # library -- bit-identical to a one-GPU run, which is MEASURED below against a one-GPU run of the same rows.
    from reductive_b200.dist import Comm, ShardedKMeans, cuda_finalize, cuda_local_step, shard_rows

    n3, M3, dsub3, iters3 = 1_000_000, 96, 8, 25

    def block3(r):
        a, b = shard_rows(n3, r, world)
        gb = torch.Generator(device=dev)
        gb.manual_seed(77 + r)
        return torch.randn((b - a, M3 * dsub3), generator=gb, device=dev, dtype=torch.float32)

    lo, hi = shard_rows(n3, rank, world)
    x3 = block3(rank)
    # initial centroid (m, j) = row (7919 j + 104729 m) mod n of the training matrix, columns of m
    jj = torch.arange(K_CENTROIDS, device=dev).view(1, -1)
    mm = torch.arange(M3, device=dev).view(-1, 1)
    rows0 = (7919 * jj + 104729 * mm) % n3                                   # [M, k]
    mine = (rows0 >= lo) & (rows0 < hi)
    cen0 = torch.zeros((M3, K_CENTROIDS, dsub3), device=dev, dtype=torch.float32)
    cols = (mm * dsub3).unsqueeze(-1) + torch.arange(dsub3, device=dev).view(1, 1, -1)  # [M, 1, dsub]
    picked = x3[(rows0 - lo).clamp(0, hi - lo - 1).unsqueeze(-1), cols.expand(M3, K_CENTROIDS, dsub3)]
    cen0[mine] = picked[mine]
    if world > 1:
        dist.all_reduce(cen0, op=dist.ReduceOp.SUM)  # every entry comes from exactly one rank: exact
    loss3 = torch.zeros((M3,), device=dev)
    log("k-means: initial centroids ready")
    # the training-loop state of the library (csrc/dist.cu; a single rank exchanges nothing): both layouts of the rows,
    # FP64 sum of squares, ordered update
    comm = Comm() if world > 1 else Comm(rank=0, world=1)
    km = ShardedKMeans(comm, x3, M3, K_CENTROIDS, dsub3)
    step3 = lambda c, l=None: km.iterate(c, l)  # noqa: E731
    cen3 = cen0.clone()
    for _ in range(2):
        step3(cen3)
    cen3 = cen0.clone()
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for it in range(iters3):
        step3(cen3, loss3 if it + 1 == iters3 else None)
    k1.record(stream)
    barrier()
    tk = torch.tensor([k0.elapsed_time(k1) / iters3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    it_ms = float(tk.item())
    log(f"k-means timed: {it_ms:.3f} ms/iter")
    # one pass of assignment + one pass of the update read x twice: 2 * 4 * d bytes per row + the codes (written, read)
    it_bytes = n3 * (2 * 4 * M3 * dsub3 + 2 * M3)
    it_gbs = it_bytes / (it_ms * 1e-3) / 1e9 / world
    bit_identical = None
    peer_window = km.peer_window
    km.close()
    if world > 1:
        same_everywhere = cen3.view(torch.int32).clone()
        dist.broadcast(same_everywhere, src=0)
        agree = torch.tensor([int(torch.equal(same_everywhere, cen3.view(torch.int32)))], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if rank == 0:  # the one-GPU run of the same rows
            x_all = torch.cat([block3(r) for r in range(world)])
            ref = cen0.clone()
            packed3 = torch.empty((M3 * K_CENTROIDS * dsub3 + M3 * K_CENTROIDS + M3,), device=dev)
            for _ in range(iters3):
                cuda_local_step(x_all, ref, packed3)
                cuda_finalize(packed3, n3, ref, None)
            bit_identical = bool(torch.equal(ref.view(torch.int32), cen3.view(torch.int32))) and bool(agree.item())
            del x_all
        barrier()  # every rank is done with the communicator before any rank tears its side down
    else:  # one GPU: against the stateless entry points (sort + chain kernels on the row-major rows)
        ref = cen0.clone()
        packed3 = torch.empty((M3 * K_CENTROIDS * dsub3 + M3 * K_CENTROIDS + M3,), device=dev)
        for _ in range(iters3):
            cuda_local_step(x3, ref, packed3)
            cuda_finalize(packed3, n3, ref, None)
        bit_identical = bool(torch.equal(ref.view(torch.int32), cen3.view(torch.int32)))
    comm.close()
    extra["pq_kmeans"] = {
        "workload": f"C3: Pq k-means 1M x 768 f32, 96 x 256 centroids, {iters3} iterations from SURVEY 8d's row picks, "
                    f"rows sharded over {world} GPU(s)",
        "sec_per_iter": it_ms * 1e-3, "iters_timed": iters3,
        "mode": "sharded: assignment by rows, ordered update by subquantizers (NCCL inside the C++ library)" if world > 1
                else "one GPU: tensor-core assignment + ordered update (chunk sort + one sequential chain per cluster: the reference's summation order)",
        "bit_identical_to_1gpu": bit_identical,
        "bit_identity_checked_against": "a one-GPU run of the same rows through rb_kmeans_assign_accumulate / rb_kmeans_finalize "
                                        "(sort + chain kernels)",
        "exchange_bytes_per_iter": (n3 * M3 + 4 * M3 * K_CENTROIDS * dsub3 * world) if world > 1 else 0,
        "code_exchange": ("peer-mapped code matrix: the assignment kernels store into the owners' HBM over NVLink" if peer_window
                          else "ncclSend/Recv") if world > 1 else "none (one rank)",
        "roofline": {"bound": "hbm", "achieved": it_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s per GPU",
                     "frac": it_gbs / peaks["hbm_gbs"],
                     "algorithmic_bytes_per_iter": it_bytes, "note": "two passes over x (assign, update) + codes"}}
    del x3

    log("k-means done")
    # C5: one 12.5M x 128 shard per GPU (100M rows over 8 GPUs), M = 16; device-generated, no collective
    n5, M5, dsub5 = 12_500_000, 16, 8
    pq5 = rb.Pq(None, np.random.default_rng(5).normal(size=(M5, K_CENTROIDS, dsub5)).astype(np.float32))
    g5 = torch.Generator(device=dev)
    g5.manual_seed(500 + rank)
    x5 = torch.randn((n5, M5 * dsub5), generator=g5, device=dev, dtype=torch.float32)
    c5 = torch.empty((n5, M5), dtype=torch.uint8, device=dev)
    pq5.quantize_batch_into(x5, c5)
    barrier()
    k0.record(stream)
    for _ in range(3):
        pq5.quantize_batch_into(x5, c5)
    k1.record(stream)
    barrier()
    t5 = torch.tensor([k0.elapsed_time(k1) / 3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
    c5_gbs = n5 * (4 * M5 * dsub5 + M5) / (float(t5.item()) * 1e-3) / 1e9
    extra["c5_streaming_shard"] = {"workload": f"C5: quantize_batch 12.5M x 128 per GPU x {world} GPU(s), 16 x 256 centroids",
                                   "ms": float(t5.item()), "vectors_per_s": world * n5 / (float(t5.item()) * 1e-3),
                                   "roofline": {"bound": "hbm", "achieved": c5_gbs, "peak": peaks["hbm_gbs"],
                                                "unit": "GB/s per GPU", "frac": c5_gbs / peaks["hbm_gbs"]}}
    del x5, c5
    log("C5 done")
    # C4: projected (Opq / GaussianOpq) encode + decode, 1M x 300, M = 30; exact-order FP32 rotation + encode / gather
    n4 = 1_000_000
    r4 = np.linalg.qr(np.random.default_rng(4).normal(size=(D, D)))[0].astype(np.float32)
    pq4 = rb.Pq(np.ascontiguousarray(r4), q)
    x4 = x[:n4]
    c4 = torch.empty((n4, M), dtype=torch.uint8, device=dev)
    rec4 = torch.empty((n4, D), dtype=torch.float32, device=dev)
    pq4.quantize_batch_into(x4, c4)
    pq4.reconstruct_batch_into(c4, rec4)
    barrier()
    k0.record(stream)
    pq4.quantize_batch_into(x4, c4)
    k1.record(stream)
    barrier()
    e4 = k0.elapsed_time(k1)
    k0.record(stream)
    pq4.reconstruct_batch_into(c4, rec4)
    k1.record(stream)
    barrier()
    extra["c4_projected"] = {"workload": "C4: 1M x 300 with a 300 x 300 projection, M = 30 (per GPU); rotation on tcgen05 "
                                         "(project_tc.cu): codes bit-exact, rotated reconstruction within 1e-5",
                             "encode_ms": e4, "decode_ms": k0.elapsed_time(k1),
                             "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"],
                                          "encode_achieved": n4 * (4 * D + M) / (e4 * 1e-3) / 1e9,
                                          "encode_frac": n4 * (4 * D + M) / (e4 * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                          "decode_achieved": n4 * (4 * D + M) / (k0.elapsed_time(k1) * 1e-3) / 1e9,
                                          "decode_frac": n4 * (4 * D + M) / (k0.elapsed_time(k1) * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
    del rec4, c4

    log("C4 done")
    # the device parts of Opq training on the C4 shape -- the covariance of the PCA init
    # (linalg.rs:23-44) and one train_iteration without its d x d SVD (opq.rs:161-189)
    from reductive_b200._cabi import check as _check, lib as _lib
    cov4 = torch.empty((D, D), device=dev)
    xty4 = torch.empty((D, D), device=dev)
    r4d = torch.from_numpy(np.ascontiguousarray(r4)).to(dev)
    cen4 = torch.from_numpy(q).to(dev).clone()
    st_raw = stream.cuda_stream
    opq = {}
    for name, call in (("covariance_ms", lambda: _check(_lib.rb_covariance(x4.data_ptr(), n4, D, x4.stride(0), cov4.data_ptr(), st_raw))),
                       ("train_iteration_ms", lambda: _check(_lib.rb_opq_train_iteration(
                           x4.data_ptr(), n4, D, x4.stride(0), r4d.data_ptr(), cen4.data_ptr(), M, K_CENTROIDS, xty4.data_ptr(), st_raw)))):
        call()
        barrier()
        k0.record(stream)
        for _ in range(3):
            call()
        k1.record(stream)
        barrier()
        opq[name] = k0.elapsed_time(k1) / 3
    extra["opq_training"] = {"workload": "1M x 300 per GPU, M = 30, k = 256: rb_covariance (column means + centred Gram on tcgen05) and "
                                         "rb_opq_train_iteration (exact-order X.R, k-means step, encode, gather, X^T.Y^ on tcgen05)",
                             **opq, "covariance_tflops": 2 * n4 * D * D / (opq["covariance_ms"] * 1e-3) / 1e12}
    del cov4, xty4
    log("OPQ training legs done")
    # the caller-side quantized storage over the C2 codes (2M x 30 u8 + norms resident in HBM):
    # fused decode + dot of 8 / 64 queries against every stored row (rb_qstore_dot)
    store = rb.QuantizedArray(pq, codes, torch.rand((N_ROWS,), device=dev) + 0.5)
    dd = {}
    for nq in (8, 64):
        qs = torch.randn((nq, D), device=dev)
        sc = torch.empty((nq, N_ROWS), device=dev)
        store.dot(qs, sc)
        barrier()
        k0.record(stream)
        for _ in range(3):
            store.dot(qs, sc)
        k1.record(stream)
        barrier()
        ms = k0.elapsed_time(k1) / 3
        # four queries share one pass over the codes at this shape (their 4 x 30 x 256 table fills shared memory)
        dd[f"nq{nq}"] = {"ms": ms, "scores_per_s": world * nq * N_ROWS / (ms * 1e-3),
                         "table_lookups_per_s": world * nq * N_ROWS * M / (ms * 1e-3),
                         "hbm_bytes": (nq // 4) * N_ROWS * M + nq * N_ROWS * 4,
                         "hbm_gbs": ((nq // 4) * N_ROWS * M + nq * N_ROWS * 4) / (ms * 1e-3) / 1e9}
        del sc
    extra["decode_dot"] = {"workload": "QuantizedArray.dot: queries x (2M rows x 30 u8 codes, norms) per GPU, scores in f32; "
                                       "bound by shared-memory table lookups (30 per row and query), not by HBM", **dd}
    store.close()
    log("decode + dot done")
    if rank == 0:
        # roofline of the dominant kernel (the encode kernel is the whole step): algorithmic bytes per vector
        # = 4*d + M, against the measured HBM copy bandwidth — at the measured peaks the HBM bound
        # (0.376 ms) is the binding one for C2, the BF16 tensor bound (2*k*d flop/vector) is reported beside it.
        kern_ms = float(np.mean(step_ms))
        bytes_per_vec = 4 * D + M
        flops_per_vec = 2 * K_CENTROIDS * D
        achieved = N_ROWS * bytes_per_vec / (kern_ms * 1e-3) / 1e9
        tflops = N_ROWS * flops_per_vec / (kern_ms * 1e-3) / 1e12
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": ENCODE_DRAM_BYTES_PER_LAUNCH,
                    "traffic_source": "profiles/r2_encode_tc_final_ncu.txt (dram__bytes_read + write, one launch)",
                    "peak_source": peaks["source"],
                    "kernel_ms": kern_ms, "algorithmic_bytes_per_vector": bytes_per_vec,
                    "tensor_tflops_algorithmic": tflops, "tensor_frac_of_bf16_peak": tflops / peaks["bf16_tflops"],
                    "tensor_frac_of_tf32_half_peak": tflops / (peaks["bf16_tflops"] / 2)}

        # ---- CPU baseline: the oracle port on a bounded sample of the same workload (N = 1 only) --------
        from oracle import oracle as orc

        o = orc.get()
        cores = os.cpu_count() or 1
        sample = N_ROWS if world == 1 else 4_096  # N = 1: the whole batch (~3 s on 16 cores); N > 1: parity check only
        xs = x[:sample].cpu().numpy()
        o.quantize_batch(q, None, xs[:2048], np.uint8, n_threads=cores)
        t0 = time.perf_counter()
        want = o.quantize_batch(q, None, xs, np.uint8, n_threads=cores)
        cpu_dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        o.quantize_batch(q, None, xs[: sample // cores], np.uint8, n_threads=1)
        cpu_dt1 = time.perf_counter() - t0
        parity = bool(np.array_equal(want, codes[:sample].cpu().numpy()))
        cpu_baseline = {"value": sample / cpu_dt if world == 1 else None, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{sample} rows of rank 0's batch, {cores} threads row-sharded "
                                  f"(single thread, as the reference runs it: {sample // cores / cpu_dt1:.0f} vectors/s)",
                        "codes_match_gpu": parity}
        line = json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": N_ROWS, "l2": "inputs 2.4 GB per step >> 126 MB L2",
                       "encode_algo": os.environ.get("RB_ENCODE_ALGO", "auto"), "parallelism": f"rows x{world}"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_rows * D * 4,
                    "d2h_bytes_per_step": e2e_rows * M, "steps": e2e_steps, "codes_match_device_path": e2e_ok,
                    "buffers": "pinned host memory (the contract's e2e); the fields below repeat the call on pageable memory",
                    "pageable_value": e2e_pageable, "pageable_codes_match": e2e_pageable_ok,
                    "one_process_all_devices": one_process},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "extra": extra,
        })
        os.write(json_fd, (line + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _bind_to_gpu_cpus(local_rank: int):
    """Restrict this process to the CPUs NVML lists as local to GPU `local_rank`; returns the previous affinity set
    (None when NVML or the affinity call is unavailable — then nothing changes)."""
    try:
        import pynvml

        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & prev
        if cpus:
            os.sched_setaffinity(0, cpus)
        return prev
    except Exception:  # noqa: BLE001 - best effort, the bench runs unchanged without it
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    algo = os.environ.get("RB_ENCODE_ALGO")
    if args.impl == "reference":
        run_reference(args)
        return
    if algo:
        import reductive_b200 as rb

        rb.set_encode_algo({"auto": rb.ENCODE_AUTO, "exact": rb.ENCODE_EXACT, "tensor": rb.ENCODE_TENSOR}[algo])
    run_ours(args)


if __name__ == "__main__":
    main()
