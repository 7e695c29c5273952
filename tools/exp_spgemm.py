"""Experiment: A*A^T (cfg3) timing at a given scale; run under ncu for the per-kernel split."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
R = synth.cfg3_ratings(scale)
rh = K.to_handle(R); mh = K.transpose(rh); K.release_handle(rh)
print("M:", mh, flush=True)
for i in range(reps):
    t = time.perf_counter(); ch = K.mult_abt(mh, mh); dt = time.perf_counter() - t
    st = K.spgemm_stats(ch); K.release_handle(ch)
    print(f"mult_abt {dt*1e3:9.2f} ms  Z={st['out_nnz']}  P={st['products']}  P/Z={st['products']/max(st['out_nnz'],1):.2f}  "
          f"{st['out_nnz']/dt/1e9:.2f} Gnnz/s  {st['products']/dt/1e9:.1f} Gprod/s", flush=True)
# general mult_ab on a cfg4-like square matrix (hash paths)
if len(sys.argv) > 3:
    s4 = float(sys.argv[3])
    A = synth.cfg4_square(s4)
    ah = K.to_handle(A)
    for i in range(2):
        t = time.perf_counter(); ch = K.mult_ab(ah, ah); dt = time.perf_counter() - t
        st = K.spgemm_stats(ch); K.release_handle(ch)
        print(f"mult_ab(cfg4 x{s4}) {dt*1e3:9.2f} ms  nnz={A.nnz} Z={st['out_nnz']}  P={st['products']}  "
              f"{st['out_nnz']/dt/1e9:.2f} Gnnz/s  {st['products']/dt/1e9:.1f} Gprod/s", flush=True)
    t = time.perf_counter(); th = K.transpose(ah); dt = time.perf_counter() - t
    print(f"transpose(cfg4 x{s4}) {dt*1e3:.2f} ms  {A.nnz/dt/1e9:.2f} Gnnz/s", flush=True)
