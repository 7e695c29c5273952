"""Run mult_vec on cfg2 with a given spmv_mode a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
mode = int(sys.argv[1]); scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
A = synth.cfg2_spmv(scale); x = synth.dense_vector(A.ncols, 77, "f4")
K.set_option("spmv_mode", mode)
h = K.to_handle(A)
xd = torch.from_numpy(x).cuda(); yd = torch.zeros(A.nrows, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(5): K.mult_vec_dev(h, xd.data_ptr(), 4, yd.data_ptr(), st)
torch.cuda.synchronize()
print("done", A)
