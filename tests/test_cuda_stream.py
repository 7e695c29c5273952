"""
GPU: the slab-stream SpMV kernel (csr_b200/csrc/spmv_slab.cu: x staged in shared memory, entries re-laid
out per warp and slab), forced on through the ``spmv_mode`` option so that small inputs exercise it --
with small x slabs, few CTAs and few warps so that cells, pieces, carries and the block-to-block run
hand-over all occur -- against the golden vectors, the oracle and the CSR tile kernel.
"""

import os

import numpy as np
import pytest

from csr_b200 import CSR, synth
from oracle import oracle as orc
from util import cases, gmat, assert_values_close

pytestmark = pytest.mark.gpu

_Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))


# (slab bytes, CTAs, consumer warps): library defaults; tiny slabs + one warp per CTA (long streams, many
# cells per warp); 2 KB slabs on 7 CTAs x 3 warps
CONFIGS = {"default": (0, 0, 31), "tiny-slabs": (512, 4, 1), "mid": (2048, 7, 3)}


@pytest.fixture(params=list(CONFIGS), ids=list(CONFIGS))
def slab(kernel, request):
    sb, ctas, warps = CONFIGS[request.param]
    kernel.set_option("spmv_mode", 2)
    kernel.set_option("stream_slab_bytes", sb)
    kernel.set_option("stream_ctas", ctas)
    kernel.set_option("stream_warps", warps)
    try:
        yield kernel
    finally:
        kernel.set_option("spmv_mode", 0)
        kernel.set_option("stream_slab_bytes", 0)
        kernel.set_option("stream_ctas", 0)
        kernel.set_option("stream_warps", 31)


def _scale(A, x):
    vmax = 1.0 if A.values is None else float(np.abs(A.values).max(initial=0.0))
    return vmax * float(np.abs(x).max(initial=0.0)) * float(np.diff(A.rowptrs).max(initial=1)) ** 0.5


def _run(kernel, A, x):
    h = kernel.to_handle(A)
    try:
        y1 = kernel.mult_vec(h, x)
        y2 = kernel.mult_vec(h, x)      # second call reuses the cached plan
        xk = 4 if np.asarray(x).dtype == np.float32 else 8
        info = kernel.spmv_plan_info(h, xk)
    finally:
        kernel.release_handle(h)
    if A.nnz and A.nrows:
        assert info["kernel"] == "stream", info     # the kernel under test really ran
    assert np.array_equal(y1, y2, equal_nan=True), "slab SpMV must be deterministic"
    return y1


@pytest.mark.parametrize("name", cases(_Z, "mult_vec"))
def test_golden(slab, golden, name):
    a = gmat(golden, f"{name}.a")
    x = golden[f"{name}.x"]
    y = _run(slab, a, x)
    f4 = (a.values is not None and a.values.dtype == np.float32) or x.dtype == np.float32
    assert_values_close(y, golden[f"{name}.y"], 1e-5 if f4 else 1e-10, _scale(a, x))


@pytest.mark.parametrize("dtype,xdt,rp64", [("f4", "f4", False), ("f8", "f8", False), ("f4", "f8", True),
                                            ("f8", "f4", False), (None, "f4", False), (None, "f8", True)])
def test_powerlaw_many_slabs(slab, dtype, xdt, rp64):
    # 60k columns = 4 f32 slabs / 8 f64 slabs; rows from empty to 60k entries (split into chunks)
    A = synth.powerlaw_csr(4000, 60000, 300000, seed=21, dtype=dtype or "f4", alpha=1.0, values=dtype is not None)
    if rp64:
        A = CSR(A.nrows, A.ncols, A.nnz, A.rowptrs.astype(np.int64), A.colinds, A.values, _cast=False)
    x = synth.dense_vector(A.ncols, 22, xdt)
    y = _run(slab, A, x)
    f4 = dtype == "f4" or xdt == "f4"
    assert_values_close(y, orc.mult_vec(A, x), 1e-5 if f4 else 1e-10, _scale(A, x))


@pytest.mark.parametrize("shape,nnz,alpha,skew", [((20000, 100000), 1500000, 1.0, 1.0),
                                                   ((3000, 70001), 900000, 0.3, 1.0),      # all rows heavy, odd ncols
                                                   ((100000, 40000), 800000, 0.2, 2.0),    # all rows light, skewed columns
                                                   ((50, 200000), 400000, 0.5, 1.0),       # a few huge rows
                                                   ((30000, 17), 200000, 0.5, 1.0)])       # one tiny slab
def test_shapes(slab, shape, nnz, alpha, skew):
    A = synth.powerlaw_csr(shape[0], shape[1], nnz, seed=33, dtype="f4", alpha=alpha, col_skew=skew)
    x = synth.dense_vector(A.ncols, 34, "f4")
    y = _run(slab, A, x)
    assert_values_close(y, orc.mult_vec(A, x), 1e-5, _scale(A, x))


def test_unsorted_and_duplicate_columns(slab):
    rng = np.random.default_rng(3)
    n = 200000
    rows = rng.integers(0, 3000, n)
    cols = rng.integers(0, 50000, n)          # duplicates and arbitrary order inside rows
    A = CSR.from_coo(rows, cols, rng.normal(size=n), (3000, 50000))
    x = rng.normal(size=50000)
    y = _run(slab, A, x)
    assert_values_close(y, orc.mult_vec(A, x), 1e-10, _scale(A, x))


def test_matches_tile_kernel_and_nonfinite(slab):
    A = synth.powerlaw_csr(5000, 40000, 400000, seed=41, dtype="f8", alpha=0.9)
    x = synth.dense_vector(A.ncols, 42, "f8")
    x[7] = np.inf
    x[11] = np.nan
    y = _run(slab, A, x)
    slab.set_option("spmv_mode", 1)
    h = slab.to_handle(A)
    yt = slab.mult_vec(h, x)
    assert slab.spmv_plan_info(h, 8)["kernel"] == "tile"
    slab.release_handle(h)
    slab.set_option("spmv_mode", 2)
    bad = ~np.isfinite(yt)
    assert np.array_equal(~np.isfinite(y), bad), "inf/nan must propagate to exactly the same rows"
    assert np.array_equal(np.isnan(y), np.isnan(yt))
    assert_values_close(y[~bad], yt[~bad], 1e-10, _scale(A, np.where(np.isfinite(x), x, 0.0)))


def test_plan_survives_order_columns_and_filter(slab):
    A = synth.powerlaw_csr(2000, 30000, 150000, seed=51, dtype="f8", alpha=0.8)
    A.values[::7] = 0.0
    x = synth.dense_vector(A.ncols, 52, "f8")
    ref = orc.mult_vec(A, x)
    h = slab.to_handle(A)
    try:
        y0 = slab.mult_vec(h, x)
        slab.filter_zeros(h)                   # drops the plan with the old entries
        y1 = slab.mult_vec(h, x)
        slab.order_columns(h)
        y2 = slab.mult_vec(h, x)
    finally:
        slab.release_handle(h)
    for y in (y0, y1, y2):
        assert_values_close(y, ref, 1e-10, _scale(A, x))


def test_split_rows_and_block_handover(slab):
    "Rows far longer than a piece (4096) and than a 128-entry block; structure-only and int64 rowptrs too."
    for values, rp64 in ((True, False), (False, True)):
        A = synth.powerlaw_csr(300, 30000, 600000, seed=61, dtype="f8", alpha=1.2, values=values)
        assert int(np.diff(A.rowptrs).max()) > 3 * 4096
        if rp64:
            A = CSR(A.nrows, A.ncols, A.nnz, A.rowptrs.astype(np.int64), A.colinds, A.values, _cast=False)
        x = synth.dense_vector(A.ncols, 62, "f8")
        y = _run(slab, A, x)
        assert_values_close(y, orc.mult_vec(A, x), 1e-10, _scale(A, x))


def test_empty_rows_and_columns(slab):
    "Empty rows give exactly 0.0; x entries that no column touches may be non-finite."
    A = synth.powerlaw_csr(5000, 9000, 40000, seed=71, dtype="f4", alpha=1.0)
    rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptrs))
    keep = (A.colinds % 5 != 0) & (rows % 3 != 0)
    B = CSR.from_coo(rows[keep], A.colinds[keep], A.values[keep], (A.nrows, A.ncols))
    x = synth.dense_vector(B.ncols, 72, "f4")
    x[::5] = np.nan                      # never referenced
    y = _run(slab, B, x)
    ref = orc.mult_vec(B, x)
    assert np.isfinite(y).all()
    assert np.all(y[np.diff(B.rowptrs) == 0] == 0.0)
    assert_values_close(y, ref, 1e-5, _scale(B, np.where(np.isfinite(x), x, 0.0)))


def test_auto_mode_picks_by_cost_model(kernel):
    """auto: the slab kernel only when the matrix is large and x (re-read per SM) costs at most 1.6 x the entry stream,
    and only from the SECOND mult_vec of a handle on (one-shot handles, csr/csr.py:582, never build a plan)"""
    kernel.set_option("stream_min_nnz", 100000)
    try:
        A = synth.powerlaw_csr(20000, 1000, 400000, seed=81, dtype="f4", alpha=0.8)      # 148*4 KB << 3.2 MB
        Bm = synth.powerlaw_csr(2000, 200000, 150000, seed=82, dtype="f4", alpha=0.8)    # 148*800 KB >> 1.2 MB
        for M, want in ((A, "stream"), (Bm, "tile")):
            x = synth.dense_vector(M.ncols, 83, "f4")
            h = kernel.to_handle(M)
            y = kernel.mult_vec(h, x)
            assert kernel.spmv_plan_info(h, 4)["kernel"] == "tile"          # first call: no plan yet
            y2 = kernel.mult_vec(h, x)
            info = kernel.spmv_plan_info(h, 4)
            kernel.release_handle(h)
            assert info["kernel"] == want, info
            assert_values_close(y, orc.mult_vec(M, x), 1e-5, _scale(M, x))
            assert_values_close(y2, orc.mult_vec(M, x), 1e-5, _scale(M, x))
    finally:
        kernel.set_option("stream_min_nnz", 4000000)
