"""Time normalize_rows on the device: configs[2]'s ratings matrix (100k x 50k, 20M nnz, f64)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
R = synth.cfg3_ratings(1.0)
h = K.to_handle(R)
for kind in ("center", "unit", "center", "unit"):
    K.synchronize(); t = time.perf_counter(); vec = K.normalize_rows(h, kind); dt = time.perf_counter() - t
    b = R.nnz * 8 * 2 + (R.nrows + 1) * 4 + R.nrows * 8
    print(f"{kind:6s} {dt*1e3:7.3f} ms  {b/dt/1e9:7.1f} GB/s algorithmic (read + write the values once)", flush=True)
K.release_handle(h)
