"""
CPU, world_size 2 on the gloo backend: the host-side logic of the row-partitioned
multi-GPU layer (csr_b200/dist.py) -- nnz-balanced partition, x broadcast, padded
all-gather of y segments, replication of B, and block assembly with the rowptr
fix-up of CSR._assemble_shards (csr/csr.py:623-650).

The local compute is injected and is the ORACLE here (tests may use it); on a GPU
box the same layer runs the cuda kernel (tests/test_cuda_dist.py, bench.py).
"""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from csr_b200 import CSR, synth
        from csr_b200.dist import DistSpMV, partition_rows, replicate_csr, dist_multiply, assemble_blocks
        from oracle import oracle as orc

        # every rank builds the same global matrix, keeps its own row block
        A = synth.powerlaw_csr(700, 500, 20000, seed=5, dtype="f8", alpha=1.0)
        cuts = partition_rows(A.rowptrs, world)
        mine = A.subset_rows(cuts[rank], cuts[rank + 1])
        counts = [cuts[r + 1] - cuts[r] for r in range(world)]

        def compute(x, y):  # oracle-backed local SpMV writing into the gather segment
            y.copy_(torch.from_numpy(orc.mult_vec(mine, x.numpy())))

        ds = DistSpMV(mine, counts, x_dtype="f8", device=torch.device("cpu"), compute=compute)
        x = np.random.default_rng(9).standard_normal(A.ncols)
        if rank == 0:
            ds.set_x(x)          # only the root has x; step() broadcasts it
        ds.step()
        y = ds.result()
        ref = orc.mult_vec(A, x)
        ok_spmv = bool(np.array_equal(y, ref))

        # pipelined variant: two row chunks per rank, chunk gathers launched asynchronously
        def compute_rows(x, y, r0, r1):
            y.copy_(torch.from_numpy(orc.mult_vec(mine.subset_rows(r0, r1), x.numpy())))

        ds2 = DistSpMV(mine, counts, x_dtype="f8", device=torch.device("cpu"), compute=compute_rows, chunks=2)
        if rank == 0:
            ds2.set_x(x)
        ds2.step()
        ok_spmv = ok_spmv and bool(np.array_equal(ds2.result(), ref))

        # SpGEMM: B known on rank 0 only, replicated by broadcast; A row blocks; assemble
        B = synth.powerlaw_csr(500, 300, 6000, seed=6, dtype="f8", alpha=0.7) if rank == 0 else None
        Brep = replicate_csr(B, src=0, device=torch.device("cpu"))

        def mul(a, b, tr):
            m = orc.canonical(orc.mult_abt(a, b) if tr else orc.mult_ab(a, b))
            return CSR(m.nrows, m.ncols, m.nnz, m.rowptrs, m.colinds, m.values)

        Cloc = dist_multiply(mine, Brep, multiply=mul)
        C = assemble_blocks(Cloc, device=torch.device("cpu"))
        Bfull = synth.powerlaw_csr(500, 300, 6000, seed=6, dtype="f8", alpha=0.7)
        Cref = orc.canonical(orc.mult_ab(A, Bfull))
        ok_mm = (C.nnz == Cref.nnz and np.array_equal(C.rowptrs, Cref.rowptrs)
                 and np.array_equal(C.colinds, Cref.colinds) and np.array_equal(C.values, Cref.values))
        ok_rep = (Brep.nnz == Bfull.nnz and np.array_equal(Brep.colinds, Bfull.colinds)
                  and np.array_equal(Brep.values, Bfull.values) and Brep.rowptrs.dtype == Bfull.rowptrs.dtype)
        # A A^T tile by tile: my row block times the NEXT rank's block, which arrives by a ring exchange
        from csr_b200.dist import ring_exchange
        got = ring_exchange([torch.from_numpy(np.asarray(mine.rowptrs, np.int64)), torch.from_numpy(mine.colinds.copy()),
                             torch.from_numpy(mine.values.copy())])
        nxt = (rank + 1) % world
        other = A.subset_rows(cuts[nxt], cuts[nxt + 1])
        ok_ring = (np.array_equal(got[0].numpy(), np.asarray(other.rowptrs, np.int64)) and
                   np.array_equal(got[1].numpy(), other.colinds) and np.array_equal(got[2].numpy(), other.values))
        nb = len(got[0]) - 1
        tile = orc.canonical(orc.mult_abt(mine, CSR(nb, A.ncols, len(got[1]), got[0].numpy(), got[1].numpy(), got[2].numpy())))
        full = orc.canonical(orc.mult_abt(A, A))
        frp = np.asarray(full.rowptrs, np.int64)
        rows = np.repeat(np.arange(A.nrows), np.diff(frp))
        sel = (rows >= cuts[rank]) & (rows < cuts[rank + 1]) & (full.colinds >= cuts[nxt]) & (full.colinds < cuts[nxt + 1])
        ok_ring = ok_ring and tile.nnz == int(sel.sum()) and np.array_equal(tile.colinds, full.colinds[sel] - cuts[nxt]) \
            and np.array_equal(tile.values, full.values[sel])
        q.put((rank, ok_spmv, ok_mm, ok_rep and ok_ring, counts))
    finally:
        dist.destroy_process_group()


def test_partition_rows_balanced():
    sys.path.insert(0, ROOT)
    from csr_b200 import synth
    from csr_b200.dist import partition_rows
    A = synth.powerlaw_csr(5000, 4000, 200000, seed=1, dtype="f4", alpha=1.0)
    for n in (1, 2, 4, 8):
        cuts = partition_rows(A.rowptrs, n)
        assert cuts[0] == 0 and cuts[-1] == A.nrows and len(cuts) == n + 1
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        nnzs = np.diff(A.rowptrs[cuts])
        longest = np.diff(A.rowptrs).max()
        assert nnzs.max() <= A.nnz / n + longest      # balanced up to one row


def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok_spmv, ok_mm, ok_rep, counts in res:
        assert ok_spmv, f"rank {rank}: distributed SpMV differs from the unsharded result"
        assert ok_mm, f"rank {rank}: assembled SpGEMM differs from the unsharded result"
        assert ok_rep, f"rank {rank}: replicated B differs, or the ring-exchanged block / its A A^T tile"
        assert sum(counts) == 700


def test_spgemm_row_weights_model():
    """products per row (exact) and the products + 0.45 x expected-outputs weight used to cut A into rank blocks."""
    from csr_b200 import synth
    from csr_b200.dist import spgemm_row_weights, partition_by_weight
    A = synth.powerlaw_csr(300, 200, 6000, seed=3, dtype="f8", alpha=0.8)
    B = synth.powerlaw_csr(200, 150, 5000, seed=4, dtype="f8", alpha=0.8)
    w, p = spgemm_row_weights(A, np.diff(B.rowptrs), B.ncols)
    ref = np.array([np.diff(B.rowptrs)[A.colinds[A.rowptrs[i]:A.rowptrs[i + 1]]].sum() for i in range(A.nrows)])
    assert np.array_equal(p, ref)
    assert np.all(w >= p) and np.all(w <= p + 0.45 * np.minimum(p, B.ncols) + 1)
    cuts = partition_by_weight(w, 4)
    assert cuts[0] == 0 and cuts[-1] == A.nrows and all(cuts[i] <= cuts[i + 1] for i in range(4))
    E = synth.powerlaw_csr(10, 20, 0, seed=1, dtype="f8")
    w0, p0 = spgemm_row_weights(E, np.diff(B.rowptrs)[:20], 5)
    assert w0.sum() == 0 and p0.sum() == 0


def test_spgemm_row_weights_trailing_empty_rows():
    "the last non-empty row keeps ALL its products when empty rows follow it (ADVICE r1)"
    from csr_b200 import CSR
    from csr_b200.dist import spgemm_row_weights
    A = CSR(3, 3, 3, np.array([0, 3, 3, 3]), np.array([0, 1, 2]), np.ones(3))
    _, prod = spgemm_row_weights(A, np.array([5, 7, 11]), 100)
    assert list(prod) == [23, 0, 0]
