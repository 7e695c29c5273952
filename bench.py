#!/usr/bin/env python
"""
bench.py -- the hot path of BASELINE.json measured on B200.

Headline (one JSON line on stdout, rank 0):
    metric  "spmv_hbm_gbs"  algorithmic HBM GB/s of mult_vec (SURVEY.md 8d:
            nnz*(4+V) + (nrows+1)*R + ncols*X + nrows*8 bytes per call), whole job
    config  BASELINE.json configs[1]: synthetic power-law CSR, 1M x 1M, 100M nnz,
            float32 values, float32 x, float64 y, one such row block PER GPU (weak
            scaling: N GPUs hold an (N*1M) x (N*1M) matrix of N*100M nnz; a step is
            broadcast(x) -> local SpMV -> all-gather(y) over NVLink)
    value   matrix, x and y resident in HBM, CUDA-event time on the launching
            stream, max over ranks
    e2e     the same metric through the kernel module's public call
            ``K.mult_vec(h, x)`` with HOST (pinned) x and y: H2D of x and D2H of y
            inside the timed region, the matrix resident (that is what a handle is)
    roofline / cpu_baseline / clocks / gpu_launches / parity as the task contract asks.

Further objects in the same line (the other BASELINE.json configs, each with its own parity
statement against the oracle and its own CPU figure):
    zipf    configs[1] with popularity-skewed columns ("uniform + Zipf mix, report both")
    cfg0    configs[0]: ML-1M-shaped 6040 x 3706 float64 mult_vec, GPU vs numba
    spgemm  configs[2]: A*A^T (mult_abt) item-item, out-nnz/s, strong-scaled over the ranks
    cfg3    configs[3]: 5M x 5M / 500M nnz CSR->CSC transpose and mult_ab on a row block (N=1)
    cfg4    configs[4]: row-partitioned SpMV on a 2B-nnz matrix generated on the devices (N>1)

CPU arm.  ``cpu_baseline`` and ``--impl reference`` time the UNMODIFIED reference's numba kernel,
staged in oracle/_ref by oracle/stage_ref.py (kind "reference"): one core as shipped, and all
host threads through the reference's own ``CSR._shard_rows`` + a thread pool (the kernels are
nogil).  If the staged package or numba is missing the C port of the kernel (oracle/) is timed
instead (kind "port").
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
RTOL_F4, RTOL_F8 = 1e-5, 1e-10      # BASELINE.json north_star: value tolerances for float32 / float64 inputs


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key):
    "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/traffic.json)."
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[key]
        return e["bytes"], e["source"]
    except Exception:
        return None, None


def spmv_bytes(nnz, nrows, ncols, val_bytes, x_bytes, rp_bytes=4):
    return nnz * (4 + val_bytes) + (nrows + 1) * rp_bytes + ncols * x_bytes + nrows * 8


def csr_bytes(m):
    v = 0 if m.values is None else m.values.dtype.itemsize
    return m.nnz * (4 + v) + (m.nrows + 1) * m.rowptrs.dtype.itemsize


def workload_cfg1(nrows, nnz):
    "The one string both arms print for configs[1] (the driver compares them)."
    return f"BASELINE configs[1]: synthetic power-law CSR, {nrows}x{nrows}, {nnz} nnz per GPU, float32, mult_vec"


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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def time_best(fn, reps):
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def take_rows(m, idx):
    "Rows `idx` of an oracle Mat as a new Mat (host gather)."
    from oracle import oracle as orc
    rp = m.rowptrs.astype(np.int64)
    lens = rp[idx + 1] - rp[idx]
    nrp = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(lens, out=nrp[1:])
    src = np.repeat(rp[idx] - nrp[:-1], lens) + np.arange(nrp[-1])
    return orc.Mat(len(idx), m.ncols, int(nrp[-1]), nrp, m.colinds[src], None if m.values is None else m.values[src])


# ------------------------------------------------------------ the CPU arm: numba (staged reference) or the C port
class CpuArm:
    """mult_vec / mult_ab / transpose of the reference on the host cores.  kind "reference" = the unmodified numba
    kernel from oracle/_ref; kind "port" = oracle/csr_oracle.c (when the staged package or numba is missing)."""

    def __init__(self):
        from oracle import refkernel
        self.cores = cpu_count()
        self.ref = refkernel if refkernel.available() else None
        self.kind = "reference" if self.ref else "port"
        self.why = None if self.ref else refkernel.why_not()
        self.pool = self.ref.pool(self.cores) if self.ref else None
        self.name = ("lenskit/csr numba kernel (oracle/_ref, unmodified)" if self.ref
                     else f"C port of the numba kernel (oracle/csr_oracle.c); reference unavailable: {self.why}")

    def matrix(self, m):
        from oracle import oracle as orc
        return self.ref.as_ref(m) if self.ref else orc.as_mat(m)

    def shards(self, A):
        return self.ref.shards(A, self.cores) if self.ref else None

    def mult_vec(self, A, x, parts=None):
        from oracle import oracle as orc
        if self.ref:
            return self.ref.mult_vec(A, x, parts, self.pool) if parts is not None else self.ref.mult_vec(A, x)
        return orc.mult_vec_threads(A, x, self.cores) if parts is not None else orc.mult_vec(A, x)

    def transpose(self, A):
        from oracle import oracle as orc
        return A.transpose() if self.ref else orc.transpose(A)

    def mult_ab_blocks(self, A, B):
        "A*B over row blocks of A on all host threads; returns (seconds, total out-nnz)."
        from oracle import oracle as orc
        if self.ref:
            parts = self.ref.shards(A, self.cores)
            self.ref.kernel().mult_ab(parts[0].subset_rows(0, min(parts[0].nrows, 2)), B)   # JIT outside the clock
            t0 = time.perf_counter()
            out = self.ref.mult_ab_parts(parts, B, self.pool)
            return time.perf_counter() - t0, int(sum(p.nnz for p in out))
        t0 = time.perf_counter()
        out = orc.mult_threads(A, B, self.cores)
        return time.perf_counter() - t0, int(sum(p.nnz for p in out))


def spmv_parity(A, x, y_gpu, rows=None, rtol=RTOL_F4):
    """GPU y against the oracle on rows [0, rows) of the host matrix A: every element within
    rtol * sum_k |a_ik| |x_k| (the tolerance of north_star applied to the magnitude that is summed)."""
    from oracle import oracle as orc
    M = orc.as_mat(A)
    n = M.nrows if rows is None else min(rows, M.nrows)
    S = orc.subset_rows(M, 0, n) if n < M.nrows else M
    y_ref = orc.mult_vec(S, x)
    absS = orc.Mat(S.nrows, S.ncols, S.nnz, S.rowptrs, S.colinds, None if S.values is None else np.abs(S.values))
    bound = orc.mult_vec(absS, np.abs(x))
    err = np.abs(np.asarray(y_gpu)[:n] - y_ref)
    worst = float((err / np.maximum(bound, 1e-300)).max()) if n else 0.0
    ok = bool(np.all(err <= rtol * bound + 1e-300))
    return ok, f"GPU y == oracle y on {n} rows ({S.nnz} nnz): max |err| / sum|a||x| = {worst:.2e} (bound {rtol:g})"


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

    def allok(flag, what):
        bad = allsum(0.0 if flag else 1.0)
        assert bad == 0, f"{what}: parity failed on {int(bad)} rank(s)"

    env = dict(rank=rank, world=world, local_rank=local_rank, barrier=barrier, allmax=allmax, allsum=allsum,
               allok=allok, peak=peak, peak_src=peak_src, K=K, W=W, steps=steps)

    out, A, x_host, yn = bench_spmv_headline(args, env)
    if world == 1 and rank == 0:
        out["cpu_baseline"] = cpu_baseline_spmv(A, x_host, yn)
    del A
    torch.cuda.empty_cache()

    if world == 1:
        if args.zipf_skew > 1.0:
            out["zipf"] = bench_spmv_variant(args, env, args.zipf_skew)
        if not args.skip_cfg0:
            out["cfg0"] = bench_cfg0(env)
    if args.spgemm_scale > 0:
        sp = bench_spgemm(args, env)
        if rank == 0:
            out["spgemm"] = sp
    if world == 1 and args.cfg3_scale > 0:
        out["cfg3"] = bench_cfg3(args, env)
    if world > 1 and args.cfg4_nnz > 0:
        c4 = bench_cfg4(args, env)
        if rank == 0:
            out["cfg4"] = c4
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def spmv_kernel_name(K, h, x_item=4):
    info = K.spmv_plan_info(h, x_item)
    if info["kernel"] == "stream":
        return ("k_spmv_slab<float,float> (+k_slab_fixup)", info)
    return ("k_spmv_tile<int,float,float> (+k_spmv_fixup)", info)


def bench_spmv_headline(args, env):
    import torch
    import torch.distributed as dist
    from csr_b200 import _native
    from csr_b200.dist import DistSpMV
    K, rank, world, W, steps = env["K"], env["rank"], env["world"], env["W"], env["steps"]
    barrier, allmax, allsum = env["barrier"], env["allmax"], env["allsum"]
    peak = env["peak"]

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
    ds.set_x(x_host)
    ds.local_spmv()                  # first call of the handle: CSR tile kernel (one-shot handles never build a plan)
    torch.cuda.synchronize()
    t0p = time.perf_counter()
    ds.local_spmv()                  # second call: builds the slab plan when auto mode selects that kernel
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0p
    t_handle = time.perf_counter() - t0
    kname, pinfo = spmv_kernel_name(K, ds.handle)
    log(f"[rank {rank}] block {A}  gen {t_gen:.1f}s  to_handle+plan {t_handle:.2f}s (plan {t_plan*1e3:.0f} ms)  {kname} {pinfo}")

    local_bytes = ds.bytes_per_step(A.nnz, 4)
    total_bytes = allsum(float(local_bytes))
    graphed = False
    if world > 1 and args.graph != "off":
        graphed = allsum(1.0 if ds.capture() else 0.0) == world      # all ranks or none
        if not graphed:
            ds._graph = None
            log(f"[rank {rank}] step not captured in a CUDA graph: {getattr(ds, 'graph_error', '?')}")

    # ---- value: device-resident, whole step (broadcast + SpMV + all-gather)
    for _ in range(W):
        ds.step()
    barrier()
    n0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(env["local_rank"]) as clk:
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

    # ---- N>1: what the collectives cost on their own (same buffers, same stream)
    coll = None
    if world > 1:
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
        if ds.nvls is not None:
            st = torch.cuda.current_stream().cuda_stream
            nb = int(ds.x.numel() * ds.x.element_size())

            def bcast():
                if rank == 0:
                    K.mc_broadcast(ds.x_mc, ds.x.data_ptr(), nb, st)
                ds.nvls[1].barrier()
            t_b = timed(bcast)
            t_bar = timed(lambda: ds.nvls[0].barrier())
            coll = {"multicast_broadcast_x_plus_barrier_ms": round(t_b, 5), "broadcast_x_bytes": nb,
                    "barrier_ms": round(t_bar, 5),
                    "gather_y": "one multicast copy of the y segment after the slab kernel" if kname.startswith("k_spmv_slab")
                    else "inside the SpMV kernel (multimem.st per finished row)",
                    "note": "NVLink multicast through the NVSwitch (symmetric memory), timed alone"}
        elif ds.symm is None:
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

    # ---- parity of the timed paths, on every rank: (a) the device-resident step, every rank's rows arrived;
    # (b) the rank's own rows (device path and host-API path) against the oracle
    ds.step()
    torch.cuda.synchronize()
    y_all = ds.result()
    y_dev = y_all[rank * A.nrows:(rank + 1) * A.nrows]
    same = bool(np.array_equal(y_dev, yn))
    assert same or np.allclose(y_dev, yn, rtol=1e-12, atol=0), "device-resident and host-API results differ"
    if world > 1:
        # every OTHER rank's rows must have arrived here too: compare per-segment checksums
        mine = torch.tensor([float(np.sum(yn)), float(np.abs(yn).sum())], dtype=torch.float64, device="cuda")
        sums = torch.zeros(2 * world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(sums, mine)
        sums = sums.cpu().numpy().reshape(world, 2)
        for r in range(world):
            seg = y_all[r * A.nrows:(r + 1) * A.nrows]
            assert abs(float(np.sum(seg)) - sums[r, 0]) <= 1e-9 * sums[r, 1] + 1e-300, f"rank {r}'s rows did not arrive intact"
    # full block at N=1 (the numba check in cpu_baseline covers it again), a 20 % row sample per rank otherwise
    ok, ptxt = spmv_parity(A, x_host, y_dev, rows=None if world == 1 else max(A.nrows // 5, 1))
    env["allok"](ok, "SpMV " + ptxt)
    parity = ptxt + (f"; every rank checked its own rows; all {world} segments arrived on rank 0 (checksums)" if world > 1 else "") \
        + f"; host-API y {'bit-identical to' if same else 'within 1e-12 of'} the device-resident y"
    if h_e2e is not ds.handle:
        K.release_handle(h_e2e)

    slab = kname.startswith("k_spmv_slab")
    traffic, traffic_src = (None, None)
    if args.scale == 1.0 and args.col_skew == 1.0 and world == 1:
        traffic, traffic_src = ncu_traffic("spmv_slab_cfg1" if slab else "spmv_tile_cfg1")
    if args.traffic is not None:
        traffic, traffic_src = args.traffic, "--traffic"
    if world == 1:
        par = f"single GPU: step = the local SpMV ({kname})"
    elif ds.nvls is not None:
        par = (f"row-partitioned x{world}; step = NVLS multicast copy of x (root) + barrier + " +
               ("SpMV slab kernel + one coalesced copy of the finished y segment to the NVLink multicast address"
                if kname.startswith("k_spmv_slab") else
                "SpMV kernel storing each finished y row once through the NVLink multicast address (fused gather)") +
               " + barrier" + ("; the whole step is one CUDA graph launch" if graphed else ""))
    elif ds.symm is not None:
        par = (f"row-partitioned x{world}; step = NCCL broadcast(x) + SpMV kernel storing y rows into every rank's "
               "buffer over NVLink (fused gather) + barrier")
    else:
        par = (f"row-partitioned x{world}; step = NCCL broadcast(x) + local SpMV + NCCL all-gather(y) in {ds.chunks} "
               "row chunks, each gather overlapping the next chunk's SpMV")
    out = {
        "metric": "spmv_hbm_gbs", "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": steps,
        "warmup": W, "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 values/x, f64 accumulate", "data": "synthetic",
        "config": {
            "workload": workload_cfg1(A.nrows, A.nnz),
            "global_shape": [A.nrows * world, A.ncols], "global_nnz": A.nnz * world,
            "row_lengths": "rank-size power law alpha=1.0, mean 100, cap ncols, random row order",
            "columns": "stratified uniform" if args.col_skew == 1.0 else f"stratified, skew t^{args.col_skew}",
            "parallelism": par,
            "l2_policy": "inputs (>=0.75 GB per GPU) larger than the 126 MB L2; no flush needed",
            "bytes_per_step": int(total_bytes), "scale": args.scale,
        },
        "e2e": {"value": round(e2e_val, 2), "unit": "GB/s", "h2d_bytes_per_step": int(A.ncols * 4 * world),
                "d2h_bytes_per_step": int(A.nrows * 8 * world), "ms_per_step": round(e2e_s * 1e3, 5),
                "call": "csr_b200.kernels.cuda.mult_vec(handle, x_pinned, out=y_pinned); matrix resident",
                "to_handle_s": round(t_handle, 3), "plan_build_ms": round(t_plan * 1e3, 1)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": env["peak_src"], "kernel": kname, "kernel_ms": round(ms_kernel, 5),
                     "algorithmic_bytes": int(local_bytes), "plan": pinfo,
                     "frac_of_nominal_8000": round(achieved / 8000.0, 4)},
        "parity": parity,
        "clocks": clk.summary(),
    }
    if coll is not None:
        out["collectives"] = coll
    ds.close()
    del ds
    return out, A, x_host, yn.copy()


def cpu_baseline_spmv(A, x, y_gpu):
    "The reference kernel on this box's host cores, full matrix: one core (as shipped) and all threads."
    arm = CpuArm()
    Ar = arm.matrix(A)
    y1 = arm.mult_vec(Ar, x)         # JIT + parity
    Aabs = arm.matrix(type("M", (), dict(nrows=A.nrows, ncols=A.ncols, nnz=A.nnz, rowptrs=A.rowptrs, colinds=A.colinds,
                                         values=np.abs(A.values)))())
    bound = arm.mult_vec(Aabs, np.abs(x))
    err = np.abs(np.asarray(y_gpu) - y1)
    worst = float((err / np.maximum(bound, 1e-300)).max())
    assert np.all(err <= RTOL_F4 * bound + 1e-300), f"SpMV parity against the {arm.kind} kernel failed: {worst}"
    b = spmv_bytes(A.nnz, A.nrows, A.ncols, 4, 4)
    t1 = time_best(lambda: arm.mult_vec(Ar, x), 3)
    parts = arm.shards(Ar) if arm.ref else True
    arm.mult_vec(Ar, x, parts)
    tn = time_best(lambda: arm.mult_vec(Ar, x, parts), 5)
    return {"value": round(b / tn / 1e9, 3), "unit": "GB/s", "cores": arm.cores, "kind": arm.kind, "impl": arm.name,
            "sample": f"the full matrix ({A.nnz} nnz, {A.nrows} rows), best of 5; all threads = a thread pool over "
                      f"{'CSR._shard_rows (csr/csr.py:599-621)' if arm.ref else 'row blocks'}",
            "parity": f"GPU y == {arm.kind} y on all {A.nrows} rows: max |err| / sum|a||x| = {worst:.2e} (bound {RTOL_F4:g})",
            "serial_value": round(b / t1 / 1e9, 3), "serial_note": "1 core: the numba kernel as shipped is serial",
            "cpu": cpu_model()}


def time_spmv_dev(K, h, x_host, nrows, reps, warm=5):
    "Device-resident mult_vec timed with CUDA events on torch's current stream; returns (ms, y)."
    import torch
    xd = torch.from_numpy(np.ascontiguousarray(x_host)).cuda()
    yd = torch.zeros(nrows, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    xi = xd.element_size()
    for _ in range(warm):
        K.mult_vec_dev(h, xd.data_ptr(), xi, yd.data_ptr(), st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        K.mult_vec_dev(h, xd.data_ptr(), xi, yd.data_ptr(), st)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, yd.cpu().numpy()


def bench_spmv_variant(args, env, skew):
    "configs[1] with popularity-skewed columns (t -> t**skew): the 'Zipf' half of SURVEY 8d's column mix."
    from csr_b200 import synth
    K, peak = env["K"], env["peak"]
    A = synth.cfg2_spmv(args.scale, col_skew=skew)
    x = np.random.default_rng(77).standard_normal(A.ncols).astype(np.float32)
    h = K.to_handle(A)
    ms, y = time_spmv_dev(K, h, x, A.nrows, env["steps"])
    kname, _ = spmv_kernel_name(K, h)
    K.release_handle(h)
    ok, ptxt = spmv_parity(A, x, y, rows=max(A.nrows // 5, 1))
    assert ok, "SpMV (skewed columns) " + ptxt
    b = spmv_bytes(A.nnz, A.nrows, A.ncols, 4, 4)
    return {"workload": workload_cfg1(A.nrows, A.nnz) + f", columns skewed t^{skew} (popular low ids)",
            "value": round(b / ms / 1e6, 2), "unit": "GB/s", "ms": round(ms, 5), "frac": round(b / ms / 1e6 / peak, 4),
            "kernel": kname, "parity": ptxt}


def bench_cfg0(env):
    "configs[0]: ML-1M-shaped 6040 x 3706, 1,000,209 nnz, float64 -- GPU parity + the numba figures beside it."
    import torch
    from csr_b200 import synth
    K = env["K"]
    A = synth.cfg1_movielens(1.0)
    x = np.random.default_rng(78).standard_normal(A.ncols)
    h = K.to_handle(A)
    ms, y = time_spmv_dev(K, h, x, A.nrows, 200, warm=10)
    t_e2e = time_best(lambda: K.mult_vec(h, x), 20)
    y_api = K.mult_vec(h, x)
    K.release_handle(h)
    ok, ptxt = spmv_parity(A, x, y, rtol=RTOL_F8)
    assert ok and np.array_equal(y, y_api), "cfg0 " + ptxt
    arm = CpuArm()
    Ar = arm.matrix(A)
    yr = arm.mult_vec(Ar, x)
    assert np.allclose(y, yr, rtol=1e-9, atol=1e-9), f"cfg0 parity against the {arm.kind} kernel failed"
    t1 = time_best(lambda: arm.mult_vec(Ar, x), 20)
    parts = arm.shards(Ar) if arm.ref else True
    arm.mult_vec(Ar, x, parts)
    tn = time_best(lambda: arm.mult_vec(Ar, x, parts), 20)
    b = spmv_bytes(A.nnz, A.nrows, A.ncols, 8, 8)
    return {"workload": f"BASELINE configs[0]: ML-1M-shaped CSR {A.nrows}x{A.ncols}, {A.nnz} nnz, float64, mult_vec",
            "gpu_ms": round(ms, 5), "gpu_gbs": round(b / ms / 1e6, 1), "gpu_e2e_ms": round(t_e2e * 1e3, 4),
            "gpu_e2e_call": "K.mult_vec(handle, x_host) -> new host y",
            "note": "12 MB problem: L2-resident and launch-latency bound on the GPU; a parity / CPU-baseline config "
                    "(SURVEY 8d), not a roofline config",
            "cpu": {"kind": arm.kind, "impl": arm.name, "cores": arm.cores, "serial_ms": round(t1 * 1e3, 4),
                    "serial_gbs": round(b / t1 / 1e9, 2), "all_cores_ms": round(tn * 1e3, 4),
                    "all_cores_gbs": round(b / tn / 1e9, 2), "cpu_model": cpu_model()},
            "parity": ptxt + f"; GPU y within 1e-9 of the {arm.kind} y on all rows; host-API y bit-identical"}


def check_product_block(K, ch, b, e, ref, rtol=RTOL_F8, bound=None):
    """Rows [b, e) of the GPU result handle against the oracle product `ref` of the same rows: rowptrs and (canonically
    sorted) colinds bit for bit, values within rtol of sum|a||b| per element (or of |ref| when no bound is given)."""
    from oracle import oracle as orc
    sh = K.subset_rows(ch, b, e)
    G = K.from_handle(sh)
    K.release_handle(sh)
    R = orc.canonical(ref)
    if not (np.array_equal(np.asarray(G.rowptrs, np.int64), np.asarray(R.rowptrs, np.int64)) and np.array_equal(G.colinds, R.colinds)):
        return False, 0.0
    scale = np.abs(R.values) if bound is None else bound
    err = np.abs(G.values - R.values)
    return bool(np.all(err <= rtol * scale + 1e-300)), float((err / np.maximum(scale, 1e-300)).max(initial=0.0))


def bench_spgemm(args, env):
    """A*A^T (item-item, BASELINE configs[2]): M = ratings^T, C = mult_abt(M, M).  With N GPUs the rows of M are
    partitioned by a products + outputs cost model (strong scaling), M is replicated by NCCL broadcast and every rank
    multiplies its row block; output row blocks stay distributed.  Every rank checks sampled row blocks of ITS timed
    result against the oracle: structure bit for bit, values to rtol 1e-10."""
    import torch
    from csr_b200 import synth, CSR
    from csr_b200.dist import replicate_csr, partition_by_weight, spgemm_row_weights
    from oracle import oracle as orc
    K, rank, world, peak = env["K"], env["rank"], env["world"], env["peak"]
    allmax, allsum = env["allmax"], env["allsum"]
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
    r0, r1 = cuts[rank], cuts[rank + 1]
    mh = K.to_handle(M)
    ah = K.subset_rows(mh, r0, r1) if world > 1 else mh

    def once(keep=False):
        ch = K.mult_abt(ah, mh)
        st = K.spgemm_stats(ch)
        if keep:
            return ch, st
        K.release_handle(ch)
        return st

    once()                                # warm-up: sizes the memory pool
    reps = 3
    torch.cuda.synchronize()
    env["barrier"]()
    t0 = time.perf_counter()
    for _ in range(reps):
        st = once()
    dt = allmax((time.perf_counter() - t0) / reps)
    Z, P = int(allsum(float(st["out_nnz"]))), int(allsum(float(st["products"])))

    # ---- parity at full size: sampled row blocks of this rank's result against the oracle
    ch, _ = once(keep=True)
    Mo = orc.as_mat(M)
    Mt = orc.transpose(Mo)
    nblk = max(16 // world, 2)
    budget = 4e8 / world                              # products the oracle recomputes per rank
    frac = min(budget / max(float(prod_row[r0:r1].sum()), 1.0), 1.0)
    blen = max(int((r1 - r0) * frac / nblk), 1)
    starts = np.linspace(r0, max(r1 - blen, r0), nblk).astype(np.int64)
    worst, rows_checked, z_checked, ok = 0.0, 0, 0, True
    for b in starts:
        e = int(min(b + blen, r1))
        ref = orc.mult_ab(orc.subset_rows(Mo, int(b), e), Mt)
        good, w = check_product_block(K, ch, int(b) - r0, e - r0, ref)
        ok, worst = ok and good, max(worst, w)
        rows_checked += e - int(b)
        z_checked += ref.nnz
    K.release_handle(ch)
    env["allok"](ok, "SpGEMM")
    worst = allmax(worst)
    parity = (f"every rank: {nblk} row blocks of its timed result ({int(allsum(rows_checked))} rows, {int(allsum(z_checked))} "
              f"out-nnz in all) against the oracle: rowptrs and sorted colinds bit-exact, values max rel err {worst:.2e} "
              f"(bound {RTOL_F8:g}, no absolute term)")

    t0 = time.perf_counter()
    th = K.transpose(mh)
    K.synchronize()
    t_tr = time.perf_counter() - t0
    K.release_handle(th)
    b_algo = 2 * csr_bytes(M) + Z * 12 + (M.nrows + 1) * 4      # bytes(A)+bytes(B)+bytes(C)
    b_tr = csr_bytes(M) + M.nnz * 12 + (M.ncols + 1) * 4        # the transpose inside mult_abt (per rank)
    traffic, traffic_src = ncu_traffic("spgemm_fixed_cfg2") if (args.spgemm_scale == 1.0 and world == 1) else (None, None)
    res = {"metric": "spgemm_abt_out_nnz_per_s", "value": round(Z / dt, 1), "unit": "nnz/s", "n_gpus": world,
           "scaling": "strong",
           "workload": f"BASELINE configs[2] x{args.spgemm_scale}: M={M.nrows}x{M.ncols}, {M.nnz} nnz f64, mult_abt(M,M)",
           "out_nnz": Z, "products": P, "compression": round(P / max(Z, 1), 2), "ms": round(dt * 1e3, 3),
           "dense_path": st["dense_path"],
           "products_per_s": round(P / dt, 1), "broadcast_s": round(t_bcast, 3),
           "transpose_ms": round(t_tr * 1e3, 3), "transpose_gbs": round(b_tr / t_tr / 1e9, 1),
           "parity": parity,
           "roofline": {"bound": "hbm", "achieved": round((b_algo + world * b_tr) / dt / 1e9, 2), "peak": peak * world,
                        "unit": "GB/s", "frac": round((b_algo + world * b_tr) / dt / 1e9 / (peak * world), 4),
                        "bytes": "bytes(A)+bytes(B)+bytes(C)+transpose(B) per rank", "traffic": traffic,
                        "traffic_source": traffic_src,
                        "note": "P/Z products per output entry go through shared-memory accumulators, so the "
                                "algorithmic-bytes roofline is far from binding; products_per_s is the work rate"}}
    res["side_list"] = int(st.get("side_list", 0))
    if rank == 0 and world == 1:
        # ---- the inputs item-item similarity really sees: mean-centred and unit-normalised rows (the handle is
        # normalised on the device, CSR.normalize_rows).  Timing only; parity of these value ranges is
        # tests/test_cuda_large.py::test_fixed_point_any_value_range
        res["normalised"] = {}
        for kind in ("center", "unit"):
            nh = K.to_handle(M)
            K.normalize_rows(nh, kind)
            K.release_handle(K.mult_abt(nh, nh))
            K.synchronize()
            t0 = time.perf_counter()
            ch2 = K.mult_abt(nh, nh)
            t_n = time.perf_counter() - t0
            st2 = K.spgemm_stats(ch2)
            K.release_handle(ch2)
            K.release_handle(nh)
            res["normalised"][kind] = {"ms": round(t_n * 1e3, 3), "dense_path": st2["dense_path"],
                                       "side_list": int(st2.get("side_list", 0)), "out_nnz": int(st2["out_nnz"])}
        # ---- e2e: the CSR-level call with HOST arrays: upload, A*A^T, device-side zero filter, copy-out of C
        if not args.skip_spgemm_e2e:
            Mh = CSR(M.nrows, M.ncols, M.nnz, M.rowptrs, M.colinds, M.values)
            t0 = time.perf_counter()
            C_host = Mh.multiply(Mh, transpose=True)
            t_e2e = time.perf_counter() - t0
            assert C_host.nnz == Z
            res["e2e"] = {"value": round(Z / t_e2e, 1), "unit": "nnz/s", "ms": round(t_e2e * 1e3, 1),
                          "h2d_bytes_per_step": int(2 * csr_bytes(M)), "d2h_bytes_per_step": int(Z * 12 + (M.nrows + 1) * 4),
                          "call": "csr_b200.CSR.multiply(M, transpose=True) with host arrays: two uploads, mult_abt, device "
                                  "zero filter, copy-out of C into pageable NumPy arrays through pinned slots and host threads (one run)"}
            del C_host
        # ---- CPU: the reference kernel on a sample of A's rows, all cores
        arm = CpuArm()
        # every stride-th row of A (the rows are popularity-ordered, so a leading block would not be representative)
        stride = max(int(np.ceil(prod_row.sum() / 4e8)), 1)
        pick = np.arange(0, Mo.nrows, stride)
        S = take_rows(Mo, pick)
        Bt = arm.transpose(arm.matrix(M))       # outside the clock, like the GPU figure's own transpose_ms
        tcpu, zs = arm.mult_ab_blocks(arm.matrix(S), Bt)
        ps = int(prod_row[pick].sum())
        res["cpu_baseline"] = {"value": round(zs / tcpu, 1), "unit": "nnz/s", "cores": arm.cores, "kind": arm.kind,
                               "impl": arm.name, "products_per_s": round(ps / tcpu, 1),
                               "sample": f"every {stride}th row of A: {len(pick)} of {Mo.nrows} rows ({zs} out-nnz, "
                                         f"{ps} products), one run on {arm.cores} threads over CSR._shard_rows blocks, "
                                         "transpose excluded"}
    if world > 1:
        K.release_handle(ah)
    K.release_handle(mh)
    return res


def bench_cfg3(args, env):
    """configs[3]: 5M x 5M, 500M nnz, float64.  (a) stable CSR->CSC transpose of the whole matrix; (b) general
    mult_ab(A_block, A) on a leading row block sized so that the output fits HBM; Z and P reported; parity of both
    against the oracle on sampled rows."""
    import torch
    from csr_b200 import synth
    from oracle import oracle as orc
    K, peak = env["K"], env["peak"]
    t0 = time.perf_counter()
    A = synth.cfg4_square(args.cfg3_scale)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    ah = K.to_handle(A)
    K.synchronize()
    t_up = time.perf_counter() - t0
    log(f"cfg3: {A} gen {t_gen:.1f}s upload {t_up:.1f}s")
    # ---- transpose, timed alone
    th = K.transpose(ah)
    K.release_handle(th)
    K.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        th = K.transpose(ah)
        K.synchronize()
        ts.append(time.perf_counter() - t0)
        if len(ts) < 3:
            K.release_handle(th)
    t_tr = min(ts)
    b_tr = csr_bytes(A) + A.nnz * 12 + (A.ncols + 1) * A.rowptrs.dtype.itemsize
    # parity: the transpose of the transpose is A bit for bit (stable order, values included), and sampled columns
    # against the oracle's transpose of a row block is not local -- so check T against the oracle on a small slice
    # of T's rows recomputed from A: row j of T = the entries of column j of A in row order.
    T = None
    tth = K.transpose(th)
    sh = K.subset_rows(tth, 0, min(A.nrows, 200_000))
    back = K.from_handle(sh)
    K.release_handle(sh)
    n0 = back.nrows
    e0 = int(A.rowptrs[n0])
    rt_ok = (np.array_equal(np.asarray(back.rowptrs, np.int64), np.asarray(A.rowptrs[:n0 + 1], np.int64))
             and np.array_equal(back.colinds, A.colinds[:e0]) and np.array_equal(back.values, A.values[:e0]))
    K.release_handle(tth)
    # oracle on a sample: transpose the column slice [0, c1) of A on the host and compare with T's first c1 rows
    c1 = max(int(A.ncols * 0.002), 1)
    sh = K.subset_rows(th, 0, c1)
    sub = K.from_handle(sh)
    K.release_handle(sh)
    Mo = orc.as_mat(A)
    keep = A.colinds < c1
    rows_of = np.repeat(np.arange(A.nrows, dtype=np.int64), np.diff(np.asarray(A.rowptrs, np.int64)))[keep]
    cols_of, vals_of = A.colinds[keep], A.values[keep]
    order = np.argsort(cols_of, kind="stable")
    t_ok = (np.array_equal(sub.colinds, rows_of[order].astype(np.int32)) and np.array_equal(sub.values, vals_of[order].astype(np.float64))
            and np.array_equal(np.asarray(sub.rowptrs, np.int64), np.concatenate([[0], np.cumsum(np.bincount(cols_of, minlength=c1))])))
    assert rt_ok and t_ok, "cfg3 transpose parity failed"
    K.release_handle(th)
    # ---- mult_ab on a leading row block: products budget so that Z*12 B stays well inside HBM
    from csr_b200.dist import spgemm_row_weights
    lens = np.diff(np.asarray(A.rowptrs, np.int64))
    _, prod_row = spgemm_row_weights(A, lens, A.ncols)
    cum = np.cumsum(prod_row)
    nblk = int(np.searchsorted(cum, args.cfg3_products))
    nblk = max(min(nblk, A.nrows), 1)
    bh = K.subset_rows(ah, 0, nblk)
    ch = K.mult_ab(bh, ah)
    K.release_handle(ch)
    K.synchronize()
    t0 = time.perf_counter()
    ch = K.mult_ab(bh, ah)
    K.synchronize()
    t_mm = time.perf_counter() - t0
    st = K.spgemm_stats(ch)
    Z, P = st["out_nnz"], st["products"]
    # parity: a few leading and trailing rows of the block against the oracle
    nchk = max(min(int(nblk * 2e7 / max(P, 1)), nblk // 2), 1)
    ok, worst = True, 0.0
    for b, e in ((0, nchk), (nblk - nchk, nblk)):
        ref = orc.mult_ab(orc.subset_rows(Mo, b, e), Mo)
        good, w = check_product_block(K, ch, b, e, ref)
        ok, worst = ok and good, max(worst, w)
    assert ok, "cfg3 mult_ab parity failed"
    Ablk = orc.subset_rows(Mo, 0, nblk)
    b_mm = csr_bytes(Ablk) + csr_bytes(A) + Z * 12 + (nblk + 1) * (8 if Z > 2**31 - 1 else 4)
    K.release_handle(ch)
    K.release_handle(bh)
    K.release_handle(ah)
    # ---- CPU: the reference transpose and mult_ab on bounded samples
    arm = CpuArm()
    n_t = min(int(np.searchsorted(np.asarray(A.rowptrs, np.int64), 20_000_000)), A.nrows)
    St = arm.matrix(orc.subset_rows(Mo, 0, max(n_t, 1)))
    arm.transpose(arm.matrix(orc.subset_rows(Mo, 0, 10)))
    t_cpu_tr = time_best(lambda: arm.transpose(St), 2)
    n_m = min(int(np.searchsorted(cum, 3e8)), A.nrows)
    Sm = arm.matrix(orc.subset_rows(Mo, 0, max(n_m, 1)))
    t_cpu_mm, z_cpu = arm.mult_ab_blocks(Sm, arm.matrix(A))
    return {"workload": f"BASELINE configs[3] x{args.cfg3_scale}: power-law CSR {A.nrows}x{A.ncols}, {A.nnz} nnz, float64",
            "transpose": {"ms": round(t_tr * 1e3, 3), "gbs": round(b_tr / t_tr / 1e9, 1), "frac": round(b_tr / t_tr / 1e9 / peak, 4),
                          "nnz_per_s": round(A.nnz / t_tr, 1), "algorithmic_bytes": int(b_tr),
                          "parity": f"(A^T)^T == A bit for bit on the first {n0} rows (values included); the first {c1} rows "
                                    "of A^T equal a host stable column sort of A bit for bit",
                          "cpu": {"kind": arm.kind, "cores": 1, "nnz_per_s": round(St.nnz / t_cpu_tr, 1),
                                  "sample": f"leading row block with {St.nnz} nnz, best of 2, one core (the transpose is serial)"}},
            "mult_ab": {"rows": nblk, "of_rows": A.nrows, "out_nnz": int(Z), "products": int(P), "compression": round(P / max(Z, 1), 3),
                        "ms": round(t_mm * 1e3, 2), "out_nnz_per_s": round(Z / t_mm, 1), "products_per_s": round(P / t_mm, 1),
                        "gbs": round(b_mm / t_mm / 1e9, 1), "frac": round(b_mm / t_mm / 1e9 / peak, 4),
                        "rowptr_dtype": "int64" if Z > 2**31 - 1 else "int32",
                        "extrapolated_full_ms": round(t_mm * 1e3 * float(cum[-1]) / max(P, 1), 1),
                        "note": "mult_ab(A[:rows], A): a leading row block sized by --cfg3-products so that C (12 B per entry) "
                                "fits HBM beside A; the full product is extrapolated by products",
                        "parity": f"first and last {nchk} rows of the block against the oracle: structure bit-exact, "
                                  f"values max rel err {worst:.2e} (bound {RTOL_F8:g})",
                        "cpu": {"kind": arm.kind, "cores": arm.cores, "out_nnz_per_s": round(z_cpu / t_cpu_mm, 1),
                                "sample": f"leading {Sm.nrows} rows ({z_cpu} out-nnz), one run on {arm.cores} threads"}},
            "upload_s": round(t_up, 2)}


def device_powerlaw_block(nrows, ncols, nnz, seed, device):
    """A power-law row block generated ON THE DEVICE (SURVEY 8d cfg 5: 2B nnz do not fit comfortably in host memory):
    rank-size row lengths (alpha 1, cap ncols) in random row order, stratified uniform columns (unique, ascending),
    values uniform(0.5, 5) float32.  Returns device tensors (rowptrs int64, colinds int32, values float32)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    w = 1.0 / torch.arange(1, nrows + 1, dtype=torch.float64, device=device)
    lo, hi = 0.0, 1.0

    def total(c):
        return int(torch.clamp(torch.floor(c * w), 0, ncols).sum().item())
    while total(hi) < nnz:
        hi *= 2.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if total(mid) < nnz:
            lo = mid
        else:
            hi = mid
    lens = torch.clamp(torch.floor(lo * w), 0, ncols).to(torch.int64)
    rem = nnz - int(lens.sum().item())
    room = torch.nonzero(lens < ncols).flatten()[:rem]
    lens[room] += 1
    lens = lens[torch.randperm(nrows, generator=g, device=device)]
    rp = torch.zeros(nrows + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=rp[1:])
    total_nnz = int(rp[-1].item())
    ci = torch.empty(total_nnz, dtype=torch.int32, device=device)
    vs = torch.empty(total_nnz, dtype=torch.float32, device=device)
    chunk = 1 << 27
    r0 = 0
    while r0 < nrows:
        r1 = int(torch.searchsorted(rp, rp[r0] + chunk, right=True).item()) - 1
        r1 = min(max(r1, r0 + 1), nrows)
        e0, e1 = int(rp[r0].item()), int(rp[r1].item())
        n = e1 - e0
        if n:
            ln = lens[r0:r1]
            L = torch.repeat_interleave(ln, ln)
            k = torch.arange(n, dtype=torch.int64, device=device) - torch.repeat_interleave(rp[r0:r1] - e0, ln)
            a = (k * ncols) // L
            b = ((k + 1) * ncols) // L
            u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
            ci[e0:e1] = (a + (u * (b - a).to(torch.float64)).to(torch.int64)).to(torch.int32)
            del L, k, a, b, u
        r0 = r1
    vs.uniform_(0.5, 5.0, generator=g)
    return rp, ci, vs


def bench_cfg4(args, env):
    """configs[4]: row-partitioned SpMV on a 2B-nnz matrix (20M x 20M, mean 100) over the N GPUs of the box: every rank
    generates its row block on its device; a step is the same broadcast(x) -> SpMV -> gather(y) as the headline."""
    import torch
    from csr_b200.dist import DistSpMV
    from csr_b200.csr import CSR
    from oracle import oracle as orc
    K, rank, world = env["K"], env["rank"], env["world"]
    allmax, allsum = env["allmax"], env["allsum"]
    dev = torch.device("cuda", env["local_rank"])
    ncols = int(args.cfg4_nnz // 100)
    nrows = ncols // world
    nnz = int(args.cfg4_nnz // world)
    t0 = time.perf_counter()
    rp, ci, vs = device_powerlaw_block(nrows, ncols, nnz, 5 + 1000 * rank, dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    rp_is64 = nnz > 2**31 - 1
    rpd = rp if rp_is64 else rp.to(torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    h = K.from_device_arrays(nrows, ncols, nnz, rpd.data_ptr(), int(rp_is64), ci.data_ptr(), vs.data_ptr(), 4, st)
    torch.cuda.synchronize()
    # the sample the oracle recomputes: the leading rows holding ~2M entries
    ns = int(torch.searchsorted(rp, torch.tensor([2_000_000], device=dev)).item())
    ns = max(min(ns, nrows), 1)
    es = int(rp[ns].item())
    S = orc.Mat(ns, ncols, es, rp[:ns + 1].cpu().numpy(), ci[:es].cpu().numpy(), vs[:es].cpu().numpy())
    # the leading rows of this block (about cfg4_tile_nnz entries): this rank's operand of the A*A^T tile below
    tile = None
    if args.cfg4_tile_nnz > 0:
        mt = int(torch.searchsorted(rp, torch.tensor([int(args.cfg4_tile_nnz)], device=dev)).item())
        mt = max(min(mt, nrows), 1)
        et = int(rp[mt].item())
        tile = (mt, et, rp[:mt + 1].clone(), ci[:et].clone(), vs[:et].clone())
    del rp, rpd, ci, vs
    torch.cuda.empty_cache()
    shell = type("Block", (), dict(nrows=nrows, ncols=ncols, nnz=nnz, rowptrs=np.array([0, nnz], np.int64)))()
    ds = DistSpMV(shell, [nrows] * world, x_dtype="f4", kernel=K, nvls=args.nvls != "off", handle=h)
    x_host = np.random.default_rng(79).standard_normal(ncols).astype(np.float32)
    ds.set_x(x_host)
    steps = max(env["steps"] // 4, 10)
    ds.local_spmv()
    ds.local_spmv()
    torch.cuda.synchronize()
    graphed = args.graph != "off" and allsum(1.0 if ds.capture() else 0.0) == world
    if not graphed:
        ds._graph = None
    for _ in range(3):
        ds.step()
    env["barrier"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ds.step()
    e1.record()
    torch.cuda.synchronize()
    env["barrier"]()
    ms = allmax(e0.elapsed_time(e1) / steps)
    for _ in range(3):
        ds.local_spmv()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        ds.local_spmv()
    e1.record()
    torch.cuda.synchronize()
    ms_k = allmax(e0.elapsed_time(e1) / steps)
    local_bytes = ds.bytes_per_step(nnz, 4, 8 if rp_is64 else 4)
    total = allsum(float(local_bytes))
    ds.step()
    torch.cuda.synchronize()
    y = ds._seg(0)[:ns].cpu().numpy()
    ok, ptxt = spmv_parity(S, x_host, y)
    env["allok"](ok, "cfg4 SpMV " + ptxt)
    kname, _ = spmv_kernel_name(K, h)
    ds.close()
    aat = bench_cfg4_aat(args, env, tile, ncols, dev) if tile is not None else None
    return {**({"aat": aat} if aat is not None else {}),
            "workload": f"BASELINE configs[4]: row-partitioned SpMV, {ncols}x{ncols}, {nnz * world} nnz float32 over {world} GPUs "
                        f"({nrows} rows, {nnz} nnz per GPU), generated on the devices",
            "value": round(total / ms / 1e6, 1), "unit": "GB/s", "ms_per_step": round(ms, 5), "kernel_ms": round(ms_k, 5),
            "kernel": kname, "per_gpu_frac_of_peak": round(local_bytes / ms_k / 1e6 / env["peak"], 4),
            "x_bytes": ncols * 4, "y_bytes": ncols * 8, "gen_s": round(t_gen, 1),
            "parallelism": ("NVLS multicast of x + in-kernel multicast gather of y" + (", one CUDA graph per step" if graphed else ""))
            if ds.nvls is not None else "NCCL broadcast(x) + all-gather(y)",
            "parity": "every rank: " + ptxt}


def bench_cfg4_aat(args, env, tile, ncols, dev):
    """configs[4], second half: A*A^T on the 2B-nnz matrix.  The full product has ~2e11 entries (2.4 TB), so every rank
    forms ONE tile of it: (its leading rows) x (the leading rows of the NEXT rank's block)^T -- the operand block travels
    over NCCL (ring exchange of the three CSR arrays), mult_abt transposes it on the device and the wide-result path
    (expand / sort / compress) multiplies.  Parity without a host copy of 2e8-entry operands: (i) every row sum of the
    tile equals (A_r * colsum(A_s))_i -- two SpMVs through the library and one index_add; (ii) the first rows of the tile,
    restricted to the first columns, against the oracle product of the corresponding sub-blocks, bit for bit in
    structure."""
    import torch
    from csr_b200.dist import ring_exchange
    from oracle import oracle as orc
    K, rank, world = env["K"], env["rank"], env["world"]
    allmax, allsum = env["allmax"], env["allsum"]
    mt, et, rp_a, ci_a, vs_a = tile
    st = torch.cuda.current_stream().cuda_stream
    # ---- ring exchange: my tile operand goes to rank-1, rank+1's comes to me
    t0 = time.perf_counter()
    rp_b, ci_b, vs_b = ring_exchange([rp_a, ci_a, vs_a], shift=1)
    mb, eb = int(rp_b.numel()) - 1, int(ci_b.numel())
    torch.cuda.synchronize()
    t_x = allmax(time.perf_counter() - t0)
    res, ok, err = None, True, ""
    try:
        rpa32, rpb32 = rp_a.to(torch.int32), rp_b.to(torch.int32)
        ah = K.from_device_arrays(mt, ncols, et, rpa32.data_ptr(), 0, ci_a.data_ptr(), vs_a.data_ptr(), 4, st)
        bh = K.from_device_arrays(mb, ncols, eb, rpb32.data_ptr(), 0, ci_b.data_ptr(), vs_b.data_ptr(), 4, st)
        torch.cuda.synchronize()
        K.release_handle(K.mult_abt(ah, bh))        # warm-up: sizes the memory pool
        K.synchronize()
        t0 = time.perf_counter()
        ch = K.mult_abt(ah, bh)
        t_mm = time.perf_counter() - t0
        stt = K.spgemm_stats(ch)
        # (i) row sums
        colsum_b = torch.zeros(ncols, dtype=torch.float64, device=dev).index_add_(0, ci_b.long(), vs_b.double())
        y_c = torch.empty(mt, dtype=torch.float64, device=dev)
        y_a = torch.empty(mt, dtype=torch.float64, device=dev)
        ones = torch.ones(mb, dtype=torch.float64, device=dev)
        K.mult_vec_dev(ch, ones.data_ptr(), 8, y_c.data_ptr(), st)
        K.mult_vec_dev(ah, colsum_b.data_ptr(), 8, y_a.data_ptr(), st)
        torch.cuda.synchronize()
        rel = float(((y_c - y_a).abs() / y_a.abs().clamp_min(1e-300)).max().item())
        ok = ok and rel <= 1e-9
        # (ii) a corner of the tile against the oracle
        na = int(torch.searchsorted(rp_a, torch.tensor([20000], device=dev)).item())
        na = max(min(na, mt), 1)
        nb = max(min(200_000, mb), 1)
        ea_, eb_ = int(rp_a[na].item()), int(rp_b[nb].item())
        Ao = orc.Mat(na, ncols, ea_, rp_a[:na + 1].cpu().numpy(), ci_a[:ea_].cpu().numpy(), vs_a[:ea_].cpu().numpy())
        Bo = orc.Mat(nb, ncols, eb_, rp_b[:nb + 1].cpu().numpy(), ci_b[:eb_].cpu().numpy(), vs_b[:eb_].cpu().numpy())
        ref = orc.canonical(orc.mult_abt(Ao, Bo))
        sh = K.subset_rows(ch, 0, na)
        G = K.from_handle(sh)
        K.release_handle(sh)
        grp = np.asarray(G.rowptrs, np.int64)
        keep = G.colinds < nb
        rows = np.repeat(np.arange(na), np.diff(grp))
        cnt = np.bincount(rows[keep], minlength=na)
        same = (np.array_equal(cnt, np.diff(np.asarray(ref.rowptrs, np.int64))) and np.array_equal(G.colinds[keep], ref.colinds))
        worst = float((np.abs(G.values[keep] - ref.values) / np.maximum(np.abs(ref.values), 1e-300)).max(initial=0.0)) if same else float("inf")
        ok = ok and same and worst <= RTOL_F8
        res = (int(stt["out_nnz"]), int(stt["products"]), t_mm, rel, na, int(ref.nnz), worst)
        for hh in (ch, ah, bh):
            K.release_handle(hh)
    except Exception as e:      # (collectives below stay matched: a failing rank reports, nobody hangs)
        ok, err = False, f"{type(e).__name__}: {e}"
    n_ok = int(allsum(1.0 if ok and res is not None else 0.0))
    if n_ok != world:
        errs = err or ("parity" if res is not None else "")
        if errs:
            log(f"[rank {rank}] cfg4 A*A^T tile: {errs} {res}")
        return {"error": f"{world - n_ok} of {world} ranks failed", "detail": errs[:300] if rank == 0 else ""}
    Z, P, t_mm, rel, na, zref, worst = res
    t_all = allmax(t_mm)
    Zs, Ps = int(allsum(float(Z))), int(allsum(float(P)))
    rel, worst = allmax(rel), allmax(worst)
    return {"workload": f"BASELINE configs[4]: A*A^T on the {ncols}x{ncols} matrix, one tile per GPU: (leading {mt} rows, {et} nnz of "
                        f"rank r's block) x (leading rows of rank r+1's block)^T, float32 operands, float64 result",
            "tiles": world, "out_nnz": Zs, "products": Ps, "ms": round(t_all * 1e3, 2),
            "value": round(Zs / t_all, 1), "unit": "nnz/s", "products_per_s": round(Ps / t_all, 1),
            "exchange_ms": round(t_x * 1e3, 2),
            "exchange": "csr_b200.dist.ring_exchange: the operand blocks move one rank along a ring over NCCL (batch_isend_irecv of rowptrs, colinds, values)",
            "note": "the full product (~(nnz/ncols)^2 * ncols entries = 2e11, 2.4 TB) cannot exist; a tile is what a rank would "
                    "form at a time.  mult_abt = device transpose of the received block + the expand/sort/compress SpGEMM path",
            "parity": f"every rank: all {mt} row sums of its tile equal (A_r * colsum(A_s)) within {rel:.1e} (bound 1e-09); rows "
                      f"0..{na} x columns 0..200000 of the tile against the oracle product of those sub-blocks ({zref} entries on "
                      f"rank 0): structure bit-exact, values max rel err {worst:.2e} (bound {RTOL_F8:g})"}


# ---------------------------------------------------------------- reference
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    arm = CpuArm()
    A = make_block(0, world, args.scale, args.col_skew)
    x = np.random.default_rng(77).standard_normal(A.ncols).astype(np.float32)
    Ar = arm.matrix(A)
    parts = arm.shards(Ar) if arm.ref else True
    b = spmv_bytes(A.nnz, A.nrows, A.ncols, 4, 4)
    for _ in range(args.warmup):
        arm.mult_vec(Ar, x, parts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.mult_vec(Ar, x, parts)
    dt = (time.perf_counter() - t0) / args.steps
    val = round(b / dt / 1e9, 3)
    t1 = time_best(lambda: arm.mult_vec(Ar, x), 2)
    sample = (f"the full block ({A.nnz} nnz) per step, {arm.cores} threads over "
              f"{'CSR._shard_rows blocks (csr/csr.py:599-621)' if arm.ref else 'row blocks'}")
    spgemm = reference_spgemm(args, arm) if args.spgemm_scale > 0 else None
    emit({
        "impl": "reference", "metric": "spmv_hbm_gbs", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 values/x, f64 accumulate", "data": "synthetic",
        "config": {"workload": workload_cfg1(A.nrows, A.nnz), "scale": args.scale},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": arm.cores, "kind": arm.kind, "impl": arm.name,
                         "sample": sample, "serial_value": round(b / t1 / 1e9, 3), "cpu": cpu_model()},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        **({"spgemm": spgemm} if spgemm else {}),
    })


def reference_spgemm(args, arm):
    "configs[2] on the CPU arm: mult_ab(S, M^T) for the same every-k-th-row sample of A that cpu_baseline uses."
    from csr_b200 import synth
    from csr_b200.dist import spgemm_row_weights
    from oracle import oracle as orc
    R = synth.cfg3_ratings(args.spgemm_scale)
    Mo = orc.transpose(orc.as_mat(R))                     # M = ratings^T (items x users)
    user_len = np.bincount(Mo.colinds, minlength=Mo.ncols).astype(np.int64)
    _, prod_row = spgemm_row_weights(Mo, user_len, Mo.nrows)
    stride = max(int(np.ceil(prod_row.sum() / 4e8)), 1)
    pick = np.arange(0, Mo.nrows, stride)
    S = take_rows(Mo, pick)
    Bt = arm.transpose(arm.matrix(Mo))
    tcpu, zs = arm.mult_ab_blocks(arm.matrix(S), Bt)
    ps = int(prod_row[pick].sum())
    return {"metric": "spgemm_abt_out_nnz_per_s", "value": round(zs / tcpu, 1), "unit": "nnz/s", "cores": arm.cores,
            "kind": arm.kind, "products_per_s": round(ps / tcpu, 1),
            "workload": f"BASELINE configs[2] x{args.spgemm_scale}: M={Mo.nrows}x{Mo.ncols}, {Mo.nnz} nnz f64, mult_abt(M,M)",
            "sample": f"every {stride}th row of A: {len(pick)} of {Mo.nrows} rows ({zs} out-nnz, {ps} products), one run on "
                      f"{arm.cores} threads over CSR._shard_rows blocks, transpose excluded"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only)")
    ap.add_argument("--col-skew", type=float, default=1.0)
    ap.add_argument("--zipf-skew", type=float, default=2.0, help="N=1: column skew of the second configs[1] line (<= 1 skips it)")
    ap.add_argument("--skip-cfg0", action="store_true")
    ap.add_argument("--spgemm-scale", type=float, default=1.0, help="scale of configs[2] for the A*A^T leg (0 = skip)")
    ap.add_argument("--skip-spgemm-e2e", action="store_true")
    ap.add_argument("--cfg3-scale", type=float, default=1.0, help="N=1: scale of configs[3] (transpose + mult_ab), 0 = skip")
    ap.add_argument("--cfg3-products", type=float, default=4e9, help="products of the configs[3] mult_ab row block")
    ap.add_argument("--cfg4-nnz", type=float, default=2e9, help="N>1: total nnz of the configs[4] SpMV, 0 = skip")
    ap.add_argument("--cfg4-tile-nnz", type=float, default=2.5e8,
                    help="N>1: entries of each operand of the configs[4] A*A^T tile (leading rows of a rank's block), 0 = skip")
    ap.add_argument("--chunks", type=int, default=1, help="N>1: row chunks whose all-gathers overlap the next chunk's SpMV")
    ap.add_argument("--nvls", choices=["auto", "on", "off"], default="auto",
                    help="N>1: NVLink-multicast broadcast + in-kernel multicast gather (default when the box supports it)")
    ap.add_argument("--graph", choices=["on", "off"], default="on", help="N>1: capture the step in a CUDA graph")
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
