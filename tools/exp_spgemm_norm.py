"""configs[2] A*A^T on raw, mean-centred and unit-normalised ratings (the inputs of item-item similarity):
which dense path ran, how many products went through the side list, time with the phase trace (CSRK_TRACE=1)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["raw", "center", "unit"]
for kv in sys.argv[4:]:
    k, v = kv.split("="); K.set_option(k, int(v)); print("option", k, v)
R = synth.cfg3_ratings(scale)
rh = K.to_handle(R)
for kind in kinds:
    mh = K.transpose(rh)
    if kind != "raw":
        K.normalize_rows(mh, kind)
    for i in range(reps):
        t = time.perf_counter(); ch = K.mult_abt(mh, mh); dt = time.perf_counter() - t
        st = K.spgemm_stats(ch); K.release_handle(ch)
        print(f"{kind:7s} mult_abt {dt*1e3:9.2f} ms  Z={st['out_nnz']}  P={st['products']}  path={st['dense_path']} "
              f"side_list={st['side_list']} ({100*st['side_list']/max(st['products'],1):.2f}% of P)  {st['out_nnz']/dt/1e9:.2f} Gnnz/s", flush=True)
    K.release_handle(mh)
