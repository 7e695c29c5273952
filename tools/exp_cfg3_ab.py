"""BASELINE configs[3], mult_ab leg alone: mult_ab(A[0:rows], A) with the phase trace (CSRK_TRACE=1)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.079
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for kv in sys.argv[4:]:
    k, v = kv.split('='); K.set_option(k, int(v)); print('option', k, v)
t = time.perf_counter(); A = synth.cfg4_square(scale); print("gen", A, f"{time.perf_counter()-t:.1f}s", flush=True)
ah = K.to_handle(A)
rows = int(A.nrows * frac)
bh = K.subset_rows(ah, 0, rows)
for i in range(reps):
    t = time.perf_counter(); ch = K.mult_ab(bh, ah); dt = time.perf_counter() - t
    st = K.spgemm_stats(ch); K.release_handle(ch)
    print(f"mult_ab(A[0:{rows}], A) {dt*1e3:8.2f} ms  Z={st['out_nnz']} P={st['products']}  {st['out_nnz']/dt/1e9:.2f} Gnnz/s {st['products']/dt/1e9:.1f} Gprod/s", flush=True)
