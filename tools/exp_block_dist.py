"""torchrun: every rank runs ITS row block of the A*A^T leg; CSRK_TRACE on ranks 0 and world-1."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
if rank in (0, world - 1):
    os.environ["CSRK_TRACE"] = "1"
import numpy as np, torch, torch.distributed as dist
torch.cuda.set_device(lr)
use_nccl = os.environ.get("NO_NCCL") != "1"
if use_nccl:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    t = torch.zeros(1, device="cuda"); dist.all_reduce(t)
else:
    dist.init_process_group("gloo")
from csr_b200 import synth
from csr_b200.dist import partition_by_weight
from csr_b200.kernels import get_kernel
K = get_kernel("cuda")
R = synth.cfg3_ratings(1.0)
rh = K.to_handle(R); mh = K.transpose(rh); K.release_handle(rh); M = K.from_handle(mh)
lens = np.diff(M.rowptrs).astype(np.int64)
user_len = np.bincount(M.colinds, minlength=M.ncols).astype(np.int64)
prod_row = np.add.reduceat(user_len[M.colinds], np.minimum(M.rowptrs[:-1].astype(np.int64), max(M.nnz - 1, 0))) * (lens > 0)
cuts = partition_by_weight(prod_row, world)
ah = K.subset_rows(mh, cuts[rank], cuts[rank + 1])
for i in range(4):
    torch.cuda.synchronize(); dist.barrier()
    if rank in (0, world - 1): print(f"--- rank {rank} rep {i}", file=sys.stderr, flush=True)
    t = time.perf_counter(); ch = K.mult_abt(ah, mh); dt = time.perf_counter() - t
    t2 = time.perf_counter(); st = K.spgemm_stats(ch); K.release_handle(ch); dr = time.perf_counter() - t2
    print(f"rank {rank}/{world} rep {i}: Z={st['out_nnz']:11d} P={st['products']:11d}  mult_abt {dt*1e3:7.2f} ms  release {dr*1e3:6.2f} ms", flush=True)
dist.destroy_process_group()
