"""One GPU: the row blocks of the N-GPU A*A^T leg, one after the other (what each rank would run)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from csr_b200 import synth
from csr_b200.dist import partition_by_weight
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ranks = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(world))
R = synth.cfg3_ratings(1.0)
rh = K.to_handle(R); mh = K.transpose(rh); K.release_handle(rh); M = K.from_handle(mh)
lens = np.diff(M.rowptrs).astype(np.int64)
user_len = np.bincount(M.colinds, minlength=M.ncols).astype(np.int64)
prod_row = np.add.reduceat(user_len[M.colinds], np.minimum(M.rowptrs[:-1].astype(np.int64), max(M.nnz - 1, 0))) * (lens > 0)
cuts = partition_by_weight(prod_row, world)
for r in ranks:
    ah = K.subset_rows(mh, cuts[r], cuts[r + 1])
    for i in range(3):
        if i == 2: print(f"--- rank {r}/{world}: rows {cuts[r]}..{cuts[r+1]}", file=sys.stderr, flush=True)
        os.environ["CSRK_TRACE_ON"] = "1" if i == 2 else "0"
        t = time.perf_counter(); ch = K.mult_abt(ah, mh); dt = time.perf_counter() - t
        st = K.spgemm_stats(ch); K.release_handle(ch)
    print(f"rank {r}/{world}: rows {cuts[r+1]-cuts[r]:6d}  Z={st['out_nnz']:11d} P={st['products']:11d}  {dt*1e3:7.2f} ms", flush=True)
    K.release_handle(ah)
