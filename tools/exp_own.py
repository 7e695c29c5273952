import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
R = synth.cfg3_ratings(1.0)
rh = K.to_handle(R); mh = K.transpose(rh); K.release_handle(rh)
for nw in (16, 8, 16, 8):
    K.set_option("own_nw", nw)
    t = time.perf_counter(); ch = K.mult_abt(mh, mh); dt = time.perf_counter() - t
    st = K.spgemm_stats(ch); K.release_handle(ch)
    print(f"own_nw={nw:2d} mult_abt {dt*1e3:9.2f} ms  Z={st['out_nnz']} {st['out_nnz']/dt/1e9:.2f} Gnnz/s  {st['products']/dt/1e9:.1f} Gprod/s", flush=True)
