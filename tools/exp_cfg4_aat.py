"""configs[4] A*A^T tile on ONE GPU with the phase trace: a 2.5M x 20M / 250M-nnz block (what a rank of an 8-GPU run
owns) times its own transpose.  usage: CSRK_TRACE=1 python tools/exp_cfg4_aat.py [nnz] [name=value ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
nnz = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
for kv in sys.argv[2:]:
    k, v = kv.split("="); K.set_option(k, int(v)); print("option", k, v)
dev = torch.device("cuda", 0)
ncols = nnz // 100 * 8
nrows = ncols // 8
rp, ci, vs = bench.device_powerlaw_block(nrows, ncols, nnz, 5, dev)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
rp32 = rp.to(torch.int32)
ah = K.from_device_arrays(nrows, ncols, nnz, rp32.data_ptr(), 0, ci.data_ptr(), vs.data_ptr(), 4, st)
torch.cuda.synchronize()
for i in range(2):
    t = time.perf_counter(); ch = K.mult_abt(ah, ah); dt = time.perf_counter() - t
    s = K.spgemm_stats(ch); K.release_handle(ch)
    print(f"mult_abt tile {nrows}x{ncols} {dt*1e3:9.2f} ms  Z={s['out_nnz']} P={s['products']}  {s['products']/dt/1e9:.2f} Gprod/s", flush=True)
