import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csr_b200 import synth
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
R = synth.cfg3_ratings(1.0)
rh = K.to_handle(R); mh = K.transpose(rh); K.release_handle(rh)
for ft in (1024, 768, 1024):
    K.set_option("fix_threads", ft)
    for i in range(3):
        t = time.perf_counter(); ch = K.mult_abt(mh, mh); dt = time.perf_counter() - t
        st = K.spgemm_stats(ch); K.release_handle(ch)
    print(f"fix_threads={ft}: {dt*1e3:7.2f} ms  path={st['dense_path']}", flush=True)
