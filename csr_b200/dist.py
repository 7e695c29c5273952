"""
Row-partitioned multi-GPU SpMV and SpGEMM: one process per GPU, torch.distributed
(NCCL over NVLink 5 / NVSwitch) for the exchange steps, the cuda kernel for the
local work.

The partition is the parallel form of the reference's sequential sharding
(``CSR._shard_rows`` / ``_assemble_shards``, csr/csr.py:599-650): contiguous row
blocks of about equal nnz, ``split_k = searchsorted(rowptrs, k*nnz/N)``.

* SpMV: every rank holds its row block and a full copy of x.  A step is
  ``broadcast(x)`` from the root, the local SpMV on the same stream writing into this
  rank's segment of the gather buffer, and an NCCL all-gather of the y segments (padded
  to the longest block; there is no all-gather-v).
  ``fused=True`` replaces SpMV + all-gather by ONE kernel that stores every finished row
  into this rank's segment of every rank's buffer over NVLink (symmetric memory) plus a
  barrier.  Measured on 2 B200: 0.547 ms/step fused vs 0.452 ms/step with NCCL (the
  barrier and the 8-byte-granular peer stores cost more than a 16 MB all-gather), so it
  is opt-in.
* SpGEMM (A*B, A*B^T): B is replicated by three broadcasts (rowptrs, colinds,
  values); each rank multiplies its A block; output row blocks stay distributed,
  and ``assemble`` concatenates them on every rank with the int64 rowptr
  fix-up of ``_assemble_shards`` (csr.py:632-638).

The local compute is injected (``compute=``) so the host-side logic can be tested
with world_size 2 on the gloo backend; the default is the cuda kernel and there
is no other product path.
"""

from __future__ import annotations

import numpy as np


def partition_rows(rowptrs, nparts: int):
    """Row boundaries ``[b_0=0, ..., b_N=nrows]`` giving blocks of ~equal nnz
    (the rule of csr.py:609-614 applied at k*nnz/N)."""
    rp = np.asarray(rowptrs)
    nrows = len(rp) - 1
    nnz = int(rp[-1])
    cuts = [0]
    for k in range(1, nparts):
        s = int(np.searchsorted(rp, (nnz * k) // nparts, side="left"))
        cuts.append(min(max(s, cuts[-1]), nrows))
    cuts.append(nrows)
    return cuts


def partition_by_weight(weights, nparts: int):
    """Row boundaries giving blocks of ~equal total weight (e.g. the SpGEMM products per row)."""
    cum = np.concatenate([[0], np.cumsum(np.asarray(weights, dtype=np.int64))])
    return partition_rows(cum, nparts)


def spgemm_row_weights(a, b_row_lengths, out_cols: int):
    """Cost model for partitioning the rows of A in C = A B over ranks: products per row plus 0.45 x
    the expected number of output entries (fitted on 8 B200, configs[2]: 16.7 ps per product, 7.4 ps
    per output entry).  The output count of a row is not known before the symbolic pass; with P
    products thrown at ``out_cols`` columns it is estimated as ``out_cols * (1 - exp(-P / out_cols))``.
    ``b_row_lengths[j]`` is the length of row j of B (for A B^T: the column counts of B).
    Returns (weights, products), both int64[nrows]."""
    rp = np.asarray(a.rowptrs, dtype=np.int64)
    lens = np.diff(rp)
    bl = np.asarray(b_row_lengths, dtype=np.int64)
    if a.nnz == 0:
        return np.zeros(a.nrows, np.int64), np.zeros(a.nrows, np.int64)
    cum = np.concatenate([[0], np.cumsum(bl[a.colinds[:a.nnz]])])   # products of row i = cum[rp[i+1]] - cum[rp[i]]
    prod = cum[rp[1:]] - cum[rp[:-1]]
    z_est = out_cols * -np.expm1(-prod / max(out_cols, 1))
    return (prod + 0.45 * z_est).astype(np.int64), prod


def _dist():
    import torch.distributed as dist
    return dist


class DistSpMV:
    """y = A x with A row-partitioned over the ranks of ``group``.

    ``local`` is this rank's row block (a CSR with the GLOBAL column count).
    ``row_counts[r]`` is the number of rows rank r owns.

    ``chunks`` > 1 pipelines the step: the local rows are cut into that many nnz-balanced
    chunks, each with its own handle; the all-gather of chunk c is launched asynchronously
    (NCCL's own stream) as soon as its SpMV is enqueued, so it overlaps the SpMV of chunk
    c+1.  Every rank must use the same ``chunks``.  Measured on 8 B200 (1M x 8M block per GPU):
    0.878 ms/step with 1 chunk, 0.923 with 4, 1.068 with 8 -- the smaller kernels pay wave
    tails and share SMs with NCCL -- so the default is 1.

    ``nvls=True`` replaces both NCCL collectives by NVLink multicast through the NVSwitch
    (symmetric memory + ``multicast_ptr``): the root copies x ONCE to the multicast address of a
    symmetric x buffer (``csrk_mc_broadcast``), and the SpMV kernel stores every finished row once
    through the multicast address of its y segment (``csrk_spmv_dev_mc``) -- compute and gather are
    one kernel, and no GPU sends the same bytes world-1 times.  Two symmetric-memory barriers
    order a step (x landed / everyone entered, y landed).  This is the default on GPUs; it falls back
    to the NCCL collectives (``nvls_error`` says why) when the box has no multicast support, and
    ``fused=True`` or ``chunks>1`` select the other variants.  Measured, 1M x 8M block per GPU on 8 B200:
    0.579 ms/step against 0.880 with NCCL broadcast + all-gather (0.4265 vs 0.468 on 2).
    """

    def __init__(self, local, row_counts, *, x_dtype="f4", device=None, group=None, compute=None, kernel=None,
                 fused=False, chunks=1, nvls=True, handle=None):
        import torch
        self.torch = torch
        self.group = group
        dist = _dist()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.row_counts = [int(c) for c in row_counts]
        assert len(self.row_counts) == self.world
        assert local.nrows == self.row_counts[self.rank]
        self.nrows = sum(self.row_counts)
        self.ncols = int(local.ncols)
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        tdt = torch.float32 if np.dtype(x_dtype) == np.float32 else torch.float64
        self.x = torch.zeros(self.ncols, dtype=tdt, device=self.device)
        self.chunks = max(1, int(chunks)) if (self.world > 1 and not fused and handle is None) else 1
        # chunk c of rank r holds rows [ccut[r][c], ccut[r][c+1]) of that rank's block; every rank needs all
        # ranks' chunk sizes to strip the padding, so they are exchanged once
        # (``handle``: the block already lives on the device -- e.g. generated there -- and ``local`` only carries
        # its shape; ownership of the handle passes to this object)
        my_cuts = [0, int(local.nrows)] if handle is not None else partition_rows(local.rowptrs, self.chunks)
        self.ccut = self._exchange_cuts(my_cuts)
        self.cpad = [max(self.ccut[r][c + 1] - self.ccut[r][c] for r in range(self.world)) for c in range(self.chunks)]
        self.coff = [0]
        for c in range(self.chunks):
            self.coff.append(self.coff[-1] + self.world * self.cpad[c])
        self.pad = max(self.row_counts) if self.row_counts else 0
        # gather buffer: for every chunk, `world` segments of cpad[c] doubles
        self.ybuf = None
        self.symm = None          # symmetric-memory handle when the fused gather is active
        self.peer_ptrs = None
        if fused and compute is None and self.world > 1 and self.world <= 8 and self.device.type == "cuda":
            self._setup_fused(group)
        self.nvls = None          # (y handle, x handle) when the multicast path is active
        self.x_in = self.x        # what the kernels read (the symmetric x buffer under nvls)
        if (nvls and compute is None and self.world > 1 and self.device.type == "cuda" and self.symm is None
                and not fused and self.chunks == 1):
            self._setup_nvls(group)
        if self.ybuf is None:
            self.ybuf = torch.zeros(max(self.coff[-1], 1), dtype=torch.float64, device=self.device)
        self.handles = []
        self.handle = None
        if compute is None:
            if kernel is None:
                from .kernels import get_kernel
                kernel = get_kernel("cuda")
            self.kernel = kernel
            if handle is not None:
                self.handles = [handle]
            elif self.chunks == 1:
                self.handles = [kernel.to_handle(local)]
            else:
                self.handles = [kernel.to_handle(local.subset_rows(my_cuts[c], my_cuts[c + 1])) for c in range(self.chunks)]
            self.handle = self.handles[0]
            compute = self._cuda_compute
        else:
            self._local = local
            self._my_cuts = my_cuts
        self.compute = compute

    def _exchange_cuts(self, my_cuts):
        dist = _dist()
        if self.world == 1:
            return [list(my_cuts)]
        t = self.torch.tensor(my_cuts, dtype=self.torch.int64, device=self.device)
        out = self.torch.zeros(self.world * len(my_cuts), dtype=self.torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return [[int(v) for v in row] for row in out.cpu().numpy().reshape(self.world, len(my_cuts))]

    def _setup_fused(self, group):
        """Fused compute + collective: the gather buffer lives in symmetric memory (NVLink peer
        mappings set up by torch.distributed), and the SpMV kernel stores every finished row
        into this rank's segment of EVERY rank's buffer.  The NCCL all-gather becomes a barrier."""
        torch = self.torch
        try:
            import torch.distributed._symmetric_memory as symm_mem
            dist = _dist()
            g = group if group is not None else dist.group.WORLD
            buf = symm_mem.empty(max(self.coff[-1], 1), dtype=torch.float64, device=self.device)
            hdl = symm_mem.rendezvous(buf, g)
            buf.zero_()
            off = self.rank * self.cpad[0] * 8
            ptrs = [int(p) + off for p in hdl.buffer_ptrs]
            # local first, then the peers
            self.peer_ptrs = [ptrs[self.rank]] + [ptrs[r] for r in range(self.world) if r != self.rank]
            self.ybuf, self.symm = buf, hdl
            torch.cuda.synchronize()
            hdl.barrier()
        except Exception as e:  # no P2P / symmetric memory: keep the NCCL all-gather
            self.ybuf, self.symm, self.peer_ptrs = None, None, None
            self.fused_error = repr(e)

    def _setup_nvls(self, group):
        torch = self.torch
        try:
            import torch.distributed._symmetric_memory as symm_mem
            dist = _dist()
            g = group if group is not None else dist.group.WORLD
            ybuf = symm_mem.empty(max(self.coff[-1], 1), dtype=torch.float64, device=self.device)
            hy = symm_mem.rendezvous(ybuf, g)
            xbuf = symm_mem.empty(max(self.ncols, 4), dtype=self.x.dtype, device=self.device)
            hx = symm_mem.rendezvous(xbuf, g)
            if not getattr(hy, "multicast_ptr", 0) or not getattr(hx, "multicast_ptr", 0):
                raise RuntimeError("symmetric memory has no multicast support on this box")
            ybuf.zero_()
            xbuf.zero_()
            self.ybuf, self.x_in, self.nvls = ybuf, xbuf, (hy, hx)
            self.y_mc = int(hy.multicast_ptr) + self.rank * self.cpad[0] * 8
            self.x_mc = int(hx.multicast_ptr)
            torch.cuda.synchronize()
            hy.barrier()
        except Exception as e:  # keep the NCCL collectives
            self.ybuf, self.x_in, self.nvls = None, self.x, None
            self.nvls_error = repr(e)

    def _seg(self, c, r=None):
        "This rank's (or rank r's) segment of chunk c inside the gather buffer (padded length)."
        r = self.rank if r is None else r
        o = self.coff[c] + r * self.cpad[c]
        return self.ybuf[o:o + self.cpad[c]]

    def _cuda_compute(self, x, y, c=0):
        stream = self.torch.cuda.current_stream().cuda_stream
        x = self.x_in
        self.kernel.mult_vec_dev(self.handles[c], x.data_ptr(), x.element_size(), y.data_ptr(), stream)

    def set_x(self, x_host):
        "Load x on the root (other ranks receive it in step())."
        self.x.copy_(self.torch.as_tensor(np.ascontiguousarray(x_host)).to(self.x.dtype), non_blocking=False)
        if self.x_in is not self.x:
            # under NVLS the kernels read the symmetric buffer, which step() fills by multicast: keep the
            # local copy in step with x so that local_spmv() is valid right after set_x()
            self.x_in[:self.ncols].copy_(self.x)

    def local_spmv(self):
        "Only the local kernel(s), no collectives (bench.py times this for the roofline)."
        for c in range(self.chunks):
            n = self.ccut[self.rank][c + 1] - self.ccut[self.rank][c]
            self._run_chunk(c, self._seg(c)[:n])

    def _run_chunk(self, c, seg):
        if self.handle is not None:
            self._cuda_compute(self.x, seg, c)
        elif self.chunks == 1:
            self.compute(self.x, seg)
        else:  # injected compute (CPU tests): it is told which rows to produce
            self.compute(self.x, seg, self._my_cuts[c], self._my_cuts[c + 1])

    def capture(self, broadcast_x: bool = True) -> bool:
        """Capture one step into a CUDA graph (the multicast path only: its launches -- the copy of x, the
        symmetric-memory barriers, the SpMV kernels, the copy-out of y -- are all plain stream work); ``step()``
        then replays the graph: one launch per step instead of five to seven, so a slow host launch path cannot
        open gaps around a 0.2-0.4 ms kernel.  Every rank must call it; returns False (and keeps the eager step)
        if the capture fails -- ``graph_error`` says why."""
        self._graph = None
        if self.nvls is None or self.device.type != "cuda":
            self.graph_error = "only the NVLS path is captured"
            return False
        torch = self.torch
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):               # warm-up on the capture side: plans, pools, barrier pads
                    self._step_eager(broadcast_x)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_eager(broadcast_x)
            self._graph, self._graph_bx = g, broadcast_x
            return True
        except Exception as e:
            self._graph = None
            self.graph_error = repr(e)
            return False

    def step(self, broadcast_x: bool = True):
        """One distributed SpMV; the gather buffer then holds every rank's rows
        (``result()`` strips the padding)."""
        g = getattr(self, "_graph", None)
        if g is not None and self._graph_bx == broadcast_x:
            g.replay()
            return self.ybuf
        return self._step_eager(broadcast_x)

    def _step_eager(self, broadcast_x: bool = True):
        dist = _dist()
        if self.nvls is not None:
            hy, hx = self.nvls
            stream = self.torch.cuda.current_stream().cuda_stream
            # No rank may store rows into a peer's buffer while that peer still reads the previous
            # result: the first barrier of a step is only passed once every rank has entered the step
            # (stream order puts its readers before that).  The x barrier doubles as that barrier.
            if broadcast_x:
                if self.rank == 0:
                    self.kernel.mc_broadcast(self.x_mc, self.x.data_ptr(), self.x.numel() * self.x.element_size(), stream)
                hx.barrier()          # x has landed everywhere
            else:
                hy.barrier()
            seg = self._seg(0)
            self.kernel.mult_vec_dev_mc(self.handle, self.x_in.data_ptr(), self.x_in.element_size(), seg.data_ptr(),
                                        self.y_mc, stream)
            hy.barrier()              # all ranks' rows have landed everywhere
            return self.ybuf
        if self.world > 1 and broadcast_x:
            dist.broadcast(self.x, src=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        if self.symm is not None:
            # one kernel computes the rows and scatters them to every rank over NVLink
            stream = self.torch.cuda.current_stream().cuda_stream
            # entry barrier: no rank may store rows into a peer's buffer while that peer still reads the
            # previous result (the broadcast above orders every rank after the ROOT only)
            self.symm.barrier()
            self.kernel.mult_vec_dev_multi(self.handle, self.x.data_ptr(), self.x.element_size(), self.peer_ptrs, stream)
            self.symm.barrier()   # all ranks' segments have landed everywhere
            return self.ybuf
        works = []
        for c in range(self.chunks):
            n = self.ccut[self.rank][c + 1] - self.ccut[self.rank][c]
            self._run_chunk(c, self._seg(c)[:n])
            if self.world > 1:
                out = self.ybuf[self.coff[c]:self.coff[c + 1]]
                works.append(dist.all_gather_into_tensor(out, self._seg(c), group=self.group, async_op=self.chunks > 1))
        for w in works:
            if w is not None and self.chunks > 1:
                w.wait()          # stream-level wait: the caller's stream sees the gathered rows
        return self.ybuf

    def result(self) -> np.ndarray:
        "The assembled y on the host (padding removed), in global row order."
        y = self.ybuf.cpu().numpy()
        parts = []
        for r in range(self.world):
            for c in range(self.chunks):
                o = self.coff[c] + r * self.cpad[c]
                parts.append(y[o:o + self.ccut[r][c + 1] - self.ccut[r][c]])
        return np.concatenate(parts) if parts else np.zeros(0)

    def bytes_per_step(self, nnz_local: int, val_bytes: int, rp_bytes: int = 4) -> int:
        "Algorithmic HBM bytes of THIS rank's SpMV (SURVEY 8d): nnz*(4+V) + (rows+1)*R + ncols*X + rows*8."
        nr = self.row_counts[self.rank]
        return nnz_local * (4 + val_bytes) + (nr + 1) * rp_bytes + self.ncols * self.x.element_size() + nr * 8

    def close(self):
        for h in self.handles:
            self.kernel.release_handle(h)
        self.handles = []
        self.handle = None


def replicate_csr(mat, *, src: int = 0, group=None, device=None, csr_cls=None):
    """Broadcast a CSR from ``src`` to every rank (three NCCL broadcasts + one of the
    shape); returns the host CSR on ``src`` and a rebuilt host CSR elsewhere.
    Used to replicate B (or A^T) for the distributed SpGEMM."""
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mat
    rank = dist.get_rank(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if torch.cuda.is_available() else torch.device("cpu"))
    meta = torch.zeros(6, dtype=torch.int64, device=dev)
    if rank == src:
        vk = 0 if mat.values is None else mat.values.dtype.itemsize
        meta = torch.tensor([mat.nrows, mat.ncols, mat.nnz, mat.rowptrs.dtype.itemsize, vk, 0], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=src, group=group)
    nrows, ncols, nnz, rpw, vk, _ = (int(v) for v in meta.cpu())
    rdt = torch.int64 if rpw == 8 else torch.int32
    vdt = {0: None, 4: torch.float32, 8: torch.float64}[vk]

    def bc(host, n, dt):
        t = torch.as_tensor(np.ascontiguousarray(host)).to(dev) if rank == src else torch.empty(n, dtype=dt, device=dev)
        dist.broadcast(t, src=src, group=group)
        return t

    rp = bc(mat.rowptrs if rank == src else None, nrows + 1, rdt)
    ci = bc(mat.colinds if rank == src else None, nnz, torch.int32)
    vs = bc(mat.values if rank == src else None, nnz, vdt) if vk else None
    if rank == src:
        return mat
    if csr_cls is None:
        from .csr import CSR as csr_cls
    return csr_cls(nrows, ncols, nnz, rp.cpu().numpy(), ci.cpu().numpy(), None if vs is None else vs.cpu().numpy(),
                   _cast=False)


def ring_exchange(tensors, *, shift: int = 1, group=None):
    """Every rank sends ``tensors`` -- a list of 1-D tensors, on the CPU (gloo) or on its GPU (NCCL) -- to rank
    ``rank - shift`` and receives the list of rank ``rank + shift`` (lengths differ between ranks: they travel first,
    in one all-gather).  This is how an operand BLOCK moves when a product is formed tile by tile: C = A A^T of a
    row-partitioned A has the tiles A_r A_s^T, and a rank needs one foreign block at a time -- replicating A^T
    (``replicate_csr``) is for operands that fit N times over; configs[4] (2e9 entries) does not."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1 or shift % world == 0:
        return [t.clone() for t in tensors]
    dev = tensors[0].device
    hdr = torch.tensor([int(t.numel()) for t in tensors], dtype=torch.int64, device=dev)
    hdrs = [torch.zeros_like(hdr) for _ in range(world)]
    dist.all_gather(hdrs, hdr, group=group)
    src, dst = (rank + shift) % world, (rank - shift) % world
    recv = [torch.empty(int(n), dtype=t.dtype, device=dev) for n, t in zip(hdrs[src].tolist(), tensors)]
    ops = []
    for snd, rcv in zip(tensors, recv):
        ops.append(dist.P2POp(dist.isend, snd.contiguous(), dst, group))
        ops.append(dist.P2POp(dist.irecv, rcv, src, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return recv


def dist_multiply(a_local, b, *, transpose: bool = False, kernel=None, multiply=None):
    """This rank's block of C = A B (or A B^T): ``a_local`` is the rank's row block
    of A, ``b`` the replicated operand.  Returns the local block as a host CSR at
    kernel level (stored zeros kept).  ``multiply`` injects the local product for
    CPU tests; the default is the cuda kernel."""
    if multiply is not None:
        return multiply(a_local, b, transpose)
    if kernel is None:
        from .kernels import get_kernel
        kernel = get_kernel("cuda")
    ah, bh = kernel.to_handle(a_local), kernel.to_handle(b)
    try:
        ch = kernel.mult_abt(ah, bh) if transpose else kernel.mult_ab(ah, bh)
        try:
            return kernel.from_handle(ch)
        finally:
            kernel.release_handle(ch)
    finally:
        kernel.release_handle(ah)
        kernel.release_handle(bh)


def assemble_blocks(local, *, group=None, device=None, csr_cls=None):
    """Concatenate the ranks' output row blocks on every rank: all-gather of the
    block sizes, rowptr offset fix-up in int64 (``_assemble_shards``,
    csr.py:632-638), padded all-gather of colinds / values."""
    import torch
    dist = _dist()
    if csr_cls is None:
        from .csr import CSR as csr_cls
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if torch.cuda.is_available() else torch.device("cpu"))
    has_v = local.values is not None
    sizes = torch.tensor([local.nrows, local.nnz, local.ncols], dtype=torch.int64, device=dev)
    allsz = torch.zeros(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allsz, sizes, group=group)
    allsz = allsz.cpu().numpy().reshape(world, 3)
    max_rows, max_nnz = int(allsz[:, 0].max()), int(allsz[:, 1].max())

    def gather(host, n, pad, dt):
        t = torch.zeros(pad, dtype=dt, device=dev)
        if n:
            t[:n] = torch.as_tensor(np.ascontiguousarray(host)).to(dev).to(dt)
        out = torch.zeros(world * pad, dtype=dt, device=dev)
        dist.all_gather_into_tensor(out, t, group=group)
        return out.cpu().numpy().reshape(world, pad)

    rps = gather(np.asarray(local.rowptrs, np.int64), local.nrows + 1, max_rows + 1, torch.int64)
    cis = gather(local.colinds, local.nnz, max(max_nnz, 1), torch.int32)
    vss = gather(local.values, local.nnz, max(max_nnz, 1), torch.float64) if has_v else None
    blocks = []
    for r in range(world):
        nr, nz, nc = (int(v) for v in allsz[r])
        blocks.append(csr_cls(nr, nc, nz, rps[r, :nr + 1], cis[r, :nz], None if vss is None else vss[r, :nz]))
    return csr_cls._assemble_shards(blocks)
