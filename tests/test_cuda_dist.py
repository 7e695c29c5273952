"""
GPU (needs >= 2 GPUs on the box; skipped otherwise): the row-partitioned SpMV over NCCL with the
FUSED SpMV + gather kernel (rows stored straight into every rank's symmetric-memory buffer) and the
NVLink-multicast (NVLS) path against the NCCL all-gather path and the oracle.
"""

import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from csr_b200 import synth
        from csr_b200.dist import DistSpMV, partition_rows
        from oracle import oracle as orc
        A = synth.powerlaw_csr(30000, 20000, 900000, seed=5, dtype="f4", alpha=1.0)
        cuts = partition_rows(A.rowptrs, world)
        mine = A.subset_rows(cuts[rank], cuts[rank + 1])
        counts = [cuts[r + 1] - cuts[r] for r in range(world)]
        x = synth.dense_vector(A.ncols, 9, "f4")
        ref = orc.mult_vec(A, x)
        out = {}
        for fused in (True, False):  # fused is opt-in; the default is the NCCL all-gather
            ds = DistSpMV(mine, counts, x_dtype="f4", fused=fused, nvls=False)
            if rank == 0:
                ds.set_x(x)
            for _ in range(3):
                ds.step()
            torch.cuda.synchronize()
            dist.barrier()
            out[fused] = (ds.result(), ds.symm is not None)
            ds.close()
        yf, was_fused = out[True]
        yn, _ = out[False]
        # pipelined: 4 row chunks, each all-gather overlapping the next chunk's SpMV
        ds = DistSpMV(mine, counts, x_dtype="f4", chunks=4)
        if rank == 0:
            ds.set_x(x)
        for _ in range(3):
            ds.step()
        torch.cuda.synchronize()
        dist.barrier()
        yc = ds.result()
        ds.close()
        # NVLink multicast: x copied once to the multicast address, rows stored once by the SpMV kernel
        ds = DistSpMV(mine, counts, x_dtype="f4")  # the default
        if rank == 0:
            ds.set_x(x)
        for _ in range(3):
            ds.step()
        torch.cuda.synchronize()
        dist.barrier()
        ym, was_nvls = ds.result(), ds.nvls is not None
        # a second x through the same buffers (stale data would show)
        if rank == 0:
            ds.set_x(2.0 * x)
        ds.step()
        torch.cuda.synchronize()
        dist.barrier()
        ym2 = ds.result()
        ds.close()
        tol = dict(rtol=1e-5, atol=1e-5 * np.abs(ref).max())
        ok = bool(np.allclose(yf, ref, **tol) and np.array_equal(yf, yn) and np.allclose(yc, ref, **tol)
                  and np.array_equal(ym, yn) and np.array_equal(ym2, 2.0 * yn))
        q.put((rank, ok, (was_fused, was_nvls)))
    finally:
        dist.destroy_process_group()


def test_fused_gather_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok, ran in res:
        assert ok, f"rank {rank}: fused / multicast gather differs from the NCCL path / the oracle"
    assert all(r[2][0] for r in res), "symmetric memory was not available: the fused path did not run"
    assert all(r[2][1] for r in res), "NVLink multicast was not available: the NVLS path did not run"
