"""Microbench: NVLS x broadcast (csrk_mc_broadcast + symmetric barrier) vs NCCL broadcast."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import torch.distributed._symmetric_memory as symm_mem
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=f"cuda:{lr}")
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
src = torch.randn(nbytes // 4, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def timeit(fn, label, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{label:44s} {t.item()*1e3:8.1f} us  {nbytes/t.item()/1e6:7.1f} GB/s", flush=True)
def mc():
    if rank == 0: K.mc_broadcast(hdl.multicast_ptr, src.data_ptr(), nbytes, st)
    hdl.barrier()
timeit(mc, f"variant {os.environ.get('CSRK_MC_VARIANT','0')}: mc copy + barrier {nbytes/1e6:.0f} MB")
dist.broadcast(src, src=0); torch.cuda.synchronize()
ok = bool(torch.equal(buf, src)); 
t = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("   data correct on all ranks:", bool(t.item()), flush=True)
if os.environ.get("CSRK_MC_VARIANT", "0") == "0":
    timeit(lambda: dist.broadcast(src, src=0), "nccl broadcast")
dist.destroy_process_group()
