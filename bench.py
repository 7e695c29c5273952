#!/usr/bin/env python
"""
bench.py -- the hot path of BASELINE.json measured on B200.

Headline (one JSON line on stdout, rank 0):
    metric  "spmv_hbm_gbs"  algorithmic HBM GB/s of mult_vec (SURVEY.md 8d:
            nnz*(4+V) + (nrows+1)*R + ncols*X + nrows*8 bytes per call), whole job
    config  BASELINE.json configs[1]: synthetic power-law CSR, 1M x 1M, 100M nnz,
            float32 values, float32 x, float64 y, one such row block PER GPU (weak
            scaling: N GPUs hold an (N*1M) x (N*1M) matrix of N*100M nnz; a step is
            broadcast(x) -> local SpMV -> all-gather(y) over NCCL)
    value   matrix, x and y resident in HBM, CUDA-event time on the launching
            stream, max over ranks
    e2e     the same metric through the kernel module's public call
            ``K.mult_vec(h, x)`` with HOST (pinned) x and y: H2D of x and D2H of y
            inside the timed region, the matrix resident (that is what a handle is)
    roofline / cpu_baseline / clocks / gpu_launches as the task contract asks.
    spgemm  A*A^T (mult_abt) out-nnz/s on BASELINE.json configs[2] (item-item,
            100k users x 50k items, 20M nnz, f64), N=1 only, with its own
            roofline and CPU numbers.

``--impl reference`` times the reference's CPU implementation of the path (the
oracle port of the numba kernel; the reference itself is Python/numba and does not
exist on the GPU box) on all host threads, same metric and config, each step a
bounded row-block sample.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on
# stdout), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the
# saved original.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


PER_GPU_ROWS = 1_000_000
PER_GPU_NNZ = 100_000_000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def spmv_bytes(nnz, nrows, ncols, val_bytes, x_bytes, rp_bytes=4):
    return nnz * (4 + val_bytes) + (nrows + 1) * rp_bytes + ncols * x_bytes + nrows * 8


def csr_bytes(m):
    v = 0 if m.values is None else m.values.dtype.itemsize
    return m.nnz * (4 + v) + (m.nrows + 1) * m.rowptrs.dtype.itemsize


class ClockSampler:
    "SM clock + throttle reasons during the timed region (NVML; nvidia-smi as a fallback)."

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self._nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self._nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self._nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_block(rank, world, scale, col_skew):
    "This rank's row block of the weak-scaled cfg2 matrix (global column count)."
    from csr_b200 import synth
    nr = max(int(PER_GPU_ROWS * scale), 64)
    nnz = 100 * nr
    return synth.powerlaw_csr(nr, nr * world, nnz, seed=2 + 1000 * rank, dtype="f4", alpha=1.0, col_skew=col_skew)


def cpu_count():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def time_best(fn, reps):
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


# --------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from csr_b200 import _native
    from csr_b200.kernels import get_kernel
    from csr_b200.dist import DistSpMV
    K = get_kernel("cuda")
    W = args.warmup
    steps = args.steps
    peak, peak_src = measured_peak()

    t0 = time.perf_counter()
    A = make_block(rank, world, args.scale, args.col_skew)
    t_gen = time.perf_counter() - t0
    x_host = np.random.default_rng(77).standard_normal(A.ncols).astype(np.float32)
    row_counts = [A.nrows] * world
    t0 = time.perf_counter()
    want_nvls = world > 1 and args.nvls != "off" and not args.fused and args.chunks == 1
    if args.nvls == "on" and not want_nvls:
        raise SystemExit("--nvls on needs --gpus > 1 and neither --fused nor --chunks > 1")
    ds = DistSpMV(A, row_counts, x_dtype="f4", kernel=K, fused=args.fused, chunks=args.chunks if world > 1 else 1,
                  nvls=want_nvls)
    if want_nvls and ds.nvls is None and args.nvls == "on":
        raise SystemExit(f"--nvls on: multicast path unavailable ({getattr(ds, 'nvls_error', '?')})")
    if want_nvls and ds.nvls is None:
        log(f"[rank {rank}] NVLS path unavailable ({getattr(ds, 'nvls_error', '?')}); using the NCCL collectives")
    t_handle = time.perf_counter() - t0
    ds.set_x(x_host)
    log(f"[rank {rank}] block {A}  gen {t_gen:.1f}s  to_handle {t_handle:.2f}s")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    local_bytes = ds.bytes_per_step(A.nnz, 4)
    total_bytes = allsum(float(local_bytes))

    # ---- value: device-resident, whole step (broadcast + SpMV + all-gather)
    for _ in range(W):
        ds.step()
    barrier()
    n0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(steps):
            ds.step()
        e1.record()
        torch.cuda.synchronize()
    barrier()
    launches = _native.launch_count() - n0
    ms_step = allmax(e0.elapsed_time(e1) / steps)
    value = total_bytes / (ms_step * 1e-3) / 1e9

    # ---- the dominant kernel alone (local SpMV, no collectives) for the roofline
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ds.local_spmv()
    torch.cuda.synchronize()
    e2.record()
    for _ in range(steps):
        ds.local_spmv()
    e3.record()
    torch.cuda.synchronize()
    ms_kernel = e2.elapsed_time(e3) / steps
    achieved = local_bytes / (ms_kernel * 1e-3) / 1e9

    # ---- N>1: what the two collectives cost on their own (same buffers, same stream)
    coll = None
    if world > 1 and ds.nvls is not None:
        def timed(fn):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return allmax(a.elapsed_time(b) / steps)
        st = torch.cuda.current_stream().cuda_stream
        nb = int(ds.x.numel() * ds.x.element_size())

        def bcast():
            if rank == 0:
                K.mc_broadcast(ds.x_mc, ds.x.data_ptr(), nb, st)
            ds.nvls[1].barrier()
        t_b = timed(bcast)
        t_bar = timed(lambda: ds.nvls[0].barrier())
        coll = {"multicast_broadcast_x_plus_barrier_ms": round(t_b, 5), "broadcast_x_bytes": nb,
                "barrier_ms": round(t_bar, 5), "gather_y": "inside the SpMV kernel (multimem.st per finished row)",
                "note": "NVLink multicast through the NVSwitch (symmetric memory), timed alone"}
    elif world > 1 and ds.symm is None:
        def timed(fn):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return allmax(a.elapsed_time(b) / steps)
        t_b = timed(lambda: dist.broadcast(ds.x, src=0))
        t_g = timed(lambda: dist.all_gather_into_tensor(ds.ybuf[ds.coff[0]:ds.coff[1]], ds._seg(0)))
        coll = {"broadcast_x_ms": round(t_b, 5), "broadcast_x_bytes": int(ds.x.numel() * ds.x.element_size()),
                "all_gather_y_ms": round(t_g, 5), "all_gather_y_bytes": int((ds.coff[1] - ds.coff[0]) * 8),
                "note": "NCCL over NVLink/NVSwitch, timed alone; in a step they run back to back with the SpMV"}

    # ---- e2e: public kernel call, host (pinned) x and y
    xp = torch.empty(A.ncols, dtype=torch.float32).pin_memory()
    xp.copy_(torch.from_numpy(x_host))
    yp = torch.empty(A.nrows, dtype=torch.float64).pin_memory()
    xn, yn = xp.numpy(), yp.numpy()
    h_e2e = ds.handle if ds.chunks == 1 else K.to_handle(A)   # the public call works on one handle
    for _ in range(W):
        K.mult_vec(h_e2e, xn, out=yn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        K.mult_vec(h_e2e, xn, out=yn)
    e2e_s = allmax((time.perf_counter() - t0) / steps)
    barrier()
    e2e_val = total_bytes / e2e_s / 1e9

    # sanity: the timed path computes the right thing (size-independent check on a sample of rows)
    ds.step()
    torch.cuda.synchronize()
    y_all = ds.result()
    y_dev = y_all[rank * A.nrows:(rank + 1) * A.nrows]
    assert np.allclose(y_dev, yn, rtol=1e-9, atol=1e-9), "device-resident and host-API results differ"
    if world > 1:
        # every OTHER rank's rows must have arrived here too: compare per-segment checksums
        mine = torch.tensor([float(np.sum(yn)), float(np.abs(yn).sum())], dtype=torch.float64, device="cuda")
        sums = torch.zeros(2 * world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(sums, mine)
        sums = sums.cpu().numpy().reshape(world, 2)
        for r in range(world):
            seg = y_all[r * A.nrows:(r + 1) * A.nrows]
            assert abs(float(np.sum(seg)) - sums[r, 0]) <= 1e-9 * sums[r, 1] + 1e-300, f"rank {r}'s rows did not arrive intact"
    if h_e2e is not ds.handle:
        K.release_handle(h_e2e)

    out = {
        "metric": "spmv_hbm_gbs", "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": steps,
        "warmup": W, "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 values/x, f64 accumulate", "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[1]: synthetic power-law CSR, %dx%d, %d nnz per GPU, float32, mult_vec"
                        % (A.nrows, A.nrows, A.nnz),
            "global_shape": [A.nrows * world, A.ncols], "global_nnz": A.nnz * world,
            "row_lengths": "rank-size power law alpha=1.0, mean 100, cap ncols, random row order",
            "columns": "stratified uniform" if args.col_skew == 1.0 else f"stratified, skew t^{args.col_skew}",
            "parallelism": "single GPU: step = the local SpMV (tile kernel + carry fix-up)" if world == 1 else
                           (f"row-partitioned x{world}; step = NVLS multicast copy of x (root) + barrier + SpMV "
                            "kernel storing each finished y row once through the NVLink multicast address (fused "
                            "gather) + barrier") if ds.nvls is not None else
                           f"row-partitioned x{world}; step = NCCL broadcast(x) + " + (
                "SpMV kernel storing y rows into every rank's buffer over NVLink (fused gather) + barrier"
                if ds.symm is not None else
                f"local SpMV + NCCL all-gather(y) in {ds.chunks} row chunks, each gather overlapping the next chunk's SpMV"),
            "l2_policy": "inputs (>=0.8 GB per GPU) larger than the 126 MB L2; no flush needed",
            "bytes_per_step": int(total_bytes), "scale": args.scale,
        },
        "e2e": {"value": round(e2e_val, 2), "unit": "GB/s", "h2d_bytes_per_step": int(A.ncols * 4 * world),
                "d2h_bytes_per_step": int(A.nrows * 8 * world), "ms_per_step": round(e2e_s * 1e3, 5),
                "call": "csr_b200.kernels.cuda.mult_vec(handle, x_pinned, out=y_pinned); matrix resident",
                "to_handle_s": round(t_handle, 3)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4),
                     "traffic": args.traffic if args.traffic is not None else (
                         815749888 if (args.scale == 1.0 and args.col_skew == 1.0 and world == 1) else None),
                     "traffic_source": "profiles/r01_spmv_tile_ncu.md: dram__bytes_read.sum + dram__bytes_write.sum per launch",
                     "peak_source": peak_src,
                     "kernel": "k_spmv_tile<int,float,float> (+k_spmv_fixup, 4% of the step)", "kernel_ms": round(ms_kernel, 5),
                     "frac_of_nominal_8000": round(achieved / 8000.0, 4)},
        "clocks": clk.summary(),
    }
    if coll is not None:
        out["collectives"] = coll

    if world == 1 and rank == 0:
        out["cpu_baseline"] = cpu_baseline_spmv(A, x_host, yn)
    ds.close()
    del ds, A
    torch.cuda.empty_cache()
    if args.spgemm_scale > 0:
        sp = bench_spgemm(args, K, peak, rank, world, allmax, allsum)
        if rank == 0:
            out["spgemm"] = sp
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def cpu_baseline_spmv(A, x, y_gpu):
    "The oracle port of the numba mult_vec on this box's host cores: serial (as shipped) and all cores."
    from oracle import oracle as orc
    cores = cpu_count()
    M = orc.as_mat(A)
    # bounded sample: the leading row block holding ~20% of the nnz
    cut = int(np.searchsorted(M.rowptrs, M.nnz // 5))
    S = orc.subset_rows(M, 0, max(cut, 1))
    b = spmv_bytes(S.nnz, S.nrows, S.ncols, 4, 4)
    y_ref = orc.mult_vec(S, x)
    # parity at full size: the timed GPU path against the oracle on the sampled rows (f32 inputs: rtol 1e-5)
    scale = float(np.abs(y_ref).max())
    err = float(np.abs(y_gpu[:S.nrows] - y_ref).max())
    assert np.allclose(y_gpu[:S.nrows], y_ref, rtol=1e-5, atol=1e-5 * scale), f"SpMV parity failed: max err {err}"
    t1 = time_best(lambda: orc.mult_vec(S, x), 3)
    orc.mult_vec_threads(S, x, cores)
    tn = time_best(lambda: orc.mult_vec_threads(S, x, cores), 5)
    return {"value": round(b / tn / 1e9, 3), "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"leading row block with {S.nnz} of {M.nnz} nnz ({S.nrows} rows), best of 5",
            "parity": f"GPU y == oracle y on {S.nrows} rows (rtol 1e-5), max abs err {err:.3e}",
            "serial_value": round(b / t1 / 1e9, 3), "serial_note": "1 core: the numba kernel as shipped is serial",
            "cpu": cpu_model()}


def take_rows(m, idx):
    "Rows `idx` of an oracle Mat as a new Mat (host gather)."
    from oracle import oracle as orc
    rp = m.rowptrs.astype(np.int64)
    lens = rp[idx + 1] - rp[idx]
    nrp = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(lens, out=nrp[1:])
    src = np.repeat(rp[idx] - nrp[:-1], lens) + np.arange(nrp[-1])
    return orc.Mat(len(idx), m.ncols, int(nrp[-1]), nrp, m.colinds[src], None if m.values is None else m.values[src])


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def bench_spgemm(args, K, peak, rank, world, allmax, allsum):
    """A*A^T (item-item, BASELINE configs[2]): M = ratings^T, C = mult_abt(M, M).  With N GPUs
    the rows of M are partitioned by a products + outputs cost model (strong scaling), M is replicated by NCCL broadcast
    and every rank multiplies its row block; output row blocks stay distributed."""
    import torch
    from csr_b200 import synth
    from csr_b200.dist import replicate_csr, partition_by_weight, spgemm_row_weights
    from oracle import oracle as orc
    R = synth.cfg3_ratings(args.spgemm_scale) if rank == 0 else None
    M = None
    if rank == 0:
        rh = K.to_handle(R)
        mh0 = K.transpose(rh)             # M = ratings^T  (items x users), on the device
        K.release_handle(rh)
        M = K.from_handle(mh0)
        K.release_handle(mh0)
    t0 = time.perf_counter()
    M = replicate_csr(M, src=0)          # three NCCL broadcasts when world > 1
    t_bcast = time.perf_counter() - t0
    user_len = np.bincount(M.colinds, minlength=M.ncols).astype(np.int64)
    weight_row, prod_row = spgemm_row_weights(M, user_len, M.nrows)   # products + 0.45 x expected outputs
    cuts = partition_by_weight(weight_row, world)
    mh = K.to_handle(M)
    ah = K.subset_rows(mh, cuts[rank], cuts[rank + 1]) if world > 1 else mh

    def once():
        ch = K.mult_abt(ah, mh)
        st = K.spgemm_stats(ch)
        K.release_handle(ch)
        return st

    once()                                # warm-up: sizes the memory pool
    reps = 3
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        st = once()
    dt = allmax((time.perf_counter() - t0) / reps)
    Z, P = int(allsum(float(st["out_nnz"]))), int(allsum(float(st["products"])))
    t0 = time.perf_counter()
    th = K.transpose(mh)
    K.synchronize()
    t_tr = time.perf_counter() - t0
    K.release_handle(th)
    b_algo = 2 * csr_bytes(M) + Z * 12 + (M.nrows + 1) * 4      # bytes(A)+bytes(B)+bytes(C)
    b_tr = csr_bytes(M) + M.nnz * 12 + (M.ncols + 1) * 4        # the transpose inside mult_abt (per rank)
    res = {"metric": "spgemm_abt_out_nnz_per_s", "value": round(Z / dt, 1), "unit": "nnz/s", "n_gpus": world,
           "scaling": "strong",
           "workload": f"BASELINE configs[2] x{args.spgemm_scale}: M={M.nrows}x{M.ncols}, {M.nnz} nnz f64, mult_abt(M,M)",
           "out_nnz": Z, "products": P, "compression": round(P / max(Z, 1), 2), "ms": round(dt * 1e3, 3),
           "products_per_s": round(P / dt, 1), "broadcast_s": round(t_bcast, 3),
           "transpose_ms": round(t_tr * 1e3, 3), "transpose_gbs": round(b_tr / t_tr / 1e9, 1),
           "roofline": {"bound": "hbm", "achieved": round((b_algo + world * b_tr) / dt / 1e9, 2), "peak": peak * world,
                        "unit": "GB/s", "frac": round((b_algo + world * b_tr) / dt / 1e9 / (peak * world), 4),
                        "bytes": "bytes(A)+bytes(B)+bytes(C)+transpose(B) per rank", "traffic": None,
                        "note": "P/Z products per output entry go through shared-memory accumulators, so the "
                                "algorithmic-bytes roofline is far from binding; products_per_s is the work rate"}}
    if rank == 0 and world == 1:
        # CPU: the oracle on a leading row block sized for a few seconds, all cores
        cores = cpu_count()
        Mo = orc.as_mat(M)
        Mt = orc.transpose(Mo)
        # every stride-th row of A (the rows are popularity-ordered, so a leading block would not be representative)
        stride = max(int(np.ceil(prod_row.sum() / 4e8)), 1)
        pick = np.arange(0, Mo.nrows, stride)
        S = take_rows(Mo, pick)
        t0 = time.perf_counter()
        parts = orc.mult_threads(S, Mt, cores)
        tcpu = time.perf_counter() - t0
        zs = sum(p.nnz for p in parts)
        ps = int(prod_row[pick].sum())
        res["cpu_baseline"] = {"value": round(zs / tcpu, 1), "unit": "nnz/s", "cores": cores, "kind": "port",
                               "products_per_s": round(ps / tcpu, 1),
                               "sample": f"every {stride}th row of A: {len(pick)} of {Mo.nrows} rows ({zs} out-nnz, "
                                         f"{ps} products), one run on {cores} threads, transpose excluded"}
    if world > 1:
        K.release_handle(ah)
    K.release_handle(mh)
    return res


# ---------------------------------------------------------------- reference
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as orc
    world = int(os.environ.get("WORLD_SIZE", 1))
    cores = cpu_count()
    A = make_block(0, world, args.scale, args.col_skew)
    x = np.random.default_rng(77).standard_normal(A.ncols).astype(np.float32)
    M = orc.as_mat(A)
    cut = int(np.searchsorted(M.rowptrs, M.nnz // 5))
    S = orc.subset_rows(M, 0, max(cut, 1))
    b = spmv_bytes(S.nnz, S.nrows, S.ncols, 4, 4)
    for _ in range(args.warmup):
        orc.mult_vec_threads(S, x, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.mult_vec_threads(S, x, cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = round(b / dt / 1e9, 3)
    sample = f"leading row block with {S.nnz} of {M.nnz} nnz per step, {cores} threads over row blocks"
    emit({
        "impl": "reference", "metric": "spmv_hbm_gbs", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 values/x, f64 accumulate", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: synthetic power-law CSR, %dx%d, %d nnz, float32, mult_vec"
                               % (A.nrows, A.nrows, A.nnz), "scale": args.scale},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample,
                         "cpu": cpu_model()},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only)")
    ap.add_argument("--col-skew", type=float, default=1.0)
    ap.add_argument("--spgemm-scale", type=float, default=1.0, help="scale of configs[2] for the A*A^T leg (0 = skip)")
    ap.add_argument("--chunks", type=int, default=1, help="N>1: row chunks whose all-gathers overlap the next chunk's SpMV")
    ap.add_argument("--nvls", choices=["auto", "on", "off"], default="auto",
                    help="N>1: NVLink-multicast broadcast + in-kernel multicast gather (default when the box supports it)")
    ap.add_argument("--fused", action="store_true", help="fused SpMV+gather over peer memory instead of the NCCL all-gather")
    ap.add_argument("--traffic", type=float, default=None, help="ncu dram bytes per launch of the SpMV kernel, if known")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
