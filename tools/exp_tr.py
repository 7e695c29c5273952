import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csr_b200 import synth, _native
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
A = synth.cfg4_square(scale); ah = K.to_handle(A)
b = A.nnz * 24 + (A.nrows + A.ncols + 2) * 4
for bits in (0, 8, 9):
    K.set_option("radix_bits", bits)
    for i in range(reps):
        t = time.perf_counter(); th = K.transpose(ah); dt = time.perf_counter() - t
        K.release_handle(th)
        print(f"radix_bits={bits} transpose {dt*1e3:8.2f} ms  {b/dt/1e9:.1f} GB/s", flush=True)
