import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csr_b200 import synth, _native
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
A = synth.cfg4_square(scale); ah = K.to_handle(A)
b = A.nnz * 24 + (A.nrows + A.ncols + 2) * 4
for label, meminfo in (("plain", False), ("meminfo", True), ("plain", False)):
    for i in range(4):
        t = time.perf_counter(); th = K.transpose(ah); dt = time.perf_counter() - t
        t2 = time.perf_counter(); K.release_handle(th); dr = time.perf_counter() - t2
        extra = f" free={_native.device_info()['mem_free']/2**30:.1f} GiB" if meminfo else ""
        print(f"{label:8s} transpose {dt*1e3:8.2f} ms  release {dr*1e3:7.2f} ms  {b/dt/1e9:.1f} GB/s{extra}", flush=True)
