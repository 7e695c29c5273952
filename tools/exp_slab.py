"""cfg2 SpMV: CSR tile kernel vs slab kernel (device-resident, CUDA events), parity between the two, plan
build time.  usage: exp_slab.py [scale] [nw:piece:ring,...] [skew]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
def _cfg(c):
    v = [int(t) for t in c.split(":")]
    return tuple(v + [16, 1024, 4096, 4, 2][len(v):])   # warps : piece : ring bytes : chunks per ring : x buffers
cfgs = [_cfg(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "16").split(",")]
skew = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
t0 = time.time()
colmul = int(sys.argv[4]) if len(sys.argv) > 4 else 1    # a rank's block of the weak-scaled matrix: colmul x the columns
n = max(int(1_000_000 * scale), 64)
A = synth.powerlaw_csr(n, n * colmul, 100 * n, seed=2, dtype="f4", alpha=1.0, col_skew=skew); x = synth.dense_vector(A.ncols, 77, "f4")
print(f"gen {time.time()-t0:.1f}s {A}", flush=True)
xd = torch.from_numpy(x).cuda(); yd = torch.zeros(A.nrows, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
nbytes = A.nnz * 8 + (A.nrows + 1) * 4 + A.ncols * 4 + A.nrows * 8

def timed(h, reps=50):
    for _ in range(5): K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

K.set_option("spmv_mode", 1)
h = K.to_handle(A)
ms = timed(h); y_tile = yd.cpu().numpy().copy()
print(f"tile   {ms:.4f} ms  {nbytes/ms/1e6:.0f} GB/s  frac {nbytes/ms/1e6/6550.4:.3f}", flush=True)
K.release_handle(h)
for nw, piece, ring, nst, nxb in cfgs:
    K.set_option("spmv_mode", 2); K.set_option("stream_warps", nw); K.set_option("stream_piece", piece)
    K.set_option("stream_ring_bytes", ring); K.set_option("stream_ring_chunks", nst); K.set_option("stream_xbufs", nxb)
    h = K.to_handle(A)
    t0 = time.time(); K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st); torch.cuda.synchronize(); tb = time.time() - t0
    ms = timed(h); y = yd.cpu().numpy().copy()
    info = K.spmv_plan_info(h, 4)
    err = np.abs(y - y_tile).max() / max(np.abs(y_tile).max(), 1e-300)
    print(f"slab nw={nw} piece={piece} ring={ring}/{nst} xbufs={nxb}: {ms:.4f} ms  {nbytes/ms/1e6:.0f} GB/s  frac {nbytes/ms/1e6/6550.4:.3f}  "
          f"first call {tb*1e3:.1f} ms  max rel diff vs tile {err:.2e}  slabs {info['slabs']} x {info['slab_cols']} cols, rows/warp {info['rows_per_warp']}, stream {info['stream_bytes']/1e6:.0f} MB", flush=True)
    K.release_handle(h)
