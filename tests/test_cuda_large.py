"""
GPU: larger seeded inputs (many SpMV tiles, every SpGEMM bin, multi-pass radix
transposes) against the oracle, plus size-independent properties: row-block
independence (virtual ranks), (A^T)^T == A, linearity of mult_vec, idempotence
of order_columns.
"""

import numpy as np
import pytest

from csr_b200 import CSR, synth
from csr_b200.dist import partition_rows
from oracle import oracle as orc
from util import canonical, assert_values_close, spgemm_terms, spmv_terms

pytestmark = pytest.mark.gpu


def _mv_scale(A, x):
    "per-row sum |a||x|: the magnitude of what each y element sums (tests/util.py)"
    return spmv_terms(A, x)


@pytest.mark.parametrize("dtype,xdt,rp64", [("f4", "f4", False), ("f8", "f8", False), ("f4", "f8", True),
                                            ("f8", "f4", False), (None, "f4", False), (None, "f8", True)])
def test_spmv_powerlaw(kernel, dtype, xdt, rp64):
    # 300k nnz, rows from length 0.. to a 60k-entry row spanning 15 tiles
    A = synth.powerlaw_csr(4000, 60000, 300000, seed=21, dtype=dtype or "f4", alpha=1.0, values=dtype is not None)
    if rp64:
        A = CSR(A.nrows, A.ncols, A.nnz, A.rowptrs.astype(np.int64), A.colinds, A.values, _cast=False)
    x = synth.dense_vector(A.ncols, 22, xdt)
    h = kernel.to_handle(A)
    try:
        y = kernel.mult_vec(h, x)
        y2 = kernel.mult_vec(h, x)
    finally:
        kernel.release_handle(h)
    ref = orc.mult_vec(A, x)
    f4 = dtype == "f4" or xdt == "f4"
    assert_values_close(y, ref, 1e-5 if f4 else 1e-10, _mv_scale(A, x))
    assert np.array_equal(y, y2), "mult_vec must be deterministic (no float atomics)"


def test_spmv_empty_rows_and_tile_edges(kernel):
    # nnz an exact multiple of the 4096-entry tile, leading/trailing/inner empty rows
    rng = np.random.default_rng(5)
    lens = np.zeros(3000, np.int64)
    lens[rng.choice(3000, 1024, replace=False)] = 16      # 16384 nnz = 4 tiles exactly
    lens[0] = lens[-1] = lens[-2] = 0
    lens[1500] += 16 - lens[1500]
    lens[rng.integers(3, 2990)] += 16384 - lens.sum()
    cols = synth.stratified_columns(lens, 5000, rng)
    rp = np.zeros(3001, np.int64)
    np.cumsum(lens, out=rp[1:])
    A = CSR(3000, 5000, int(rp[-1]), rp, cols, rng.normal(size=int(rp[-1])))
    assert A.nnz % 4096 == 0
    x = rng.normal(size=5000)
    y = A.mult_vec(x)
    assert_values_close(y, orc.mult_vec(A, x), 1e-10, _mv_scale(A, x))
    assert np.all(y[lens == 0] == 0.0)


def test_spmv_pinned_output_zero_copy(kernel):
    """mult_vec(out=pinned): the kernel stores finished rows straight into the pinned host array
    (second output of the tile kernel, carries included) -- same bits as the staged copy."""
    import torch
    A = synth.powerlaw_csr(60000, 50000, 3_000_000, seed=19, dtype="f4", alpha=1.0)  # rows spanning many tiles
    x = synth.dense_vector(A.ncols, 5, "f4")
    yp = torch.full((A.nrows,), float("nan"), dtype=torch.float64).pin_memory()
    h = kernel.to_handle(A)
    try:
        y_staged = kernel.mult_vec(h, x)                      # pageable out: device buffer + D2H copy
        y_pinned = kernel.mult_vec(h, x, out=yp.numpy())      # pinned out: zero-copy stores
        kernel.set_option("spmv_zero_copy_y", 0)
        y_off = kernel.mult_vec(h, x, out=torch.empty(A.nrows, dtype=torch.float64).pin_memory().numpy()).copy()
    finally:
        kernel.set_option("spmv_zero_copy_y", 1)
        kernel.release_handle(h)
    assert np.array_equal(y_pinned, y_staged) and np.array_equal(y_off, y_staged)
    ref = orc.mult_vec(A, x)
    assert_values_close(y_pinned, ref, 1e-5, _mv_scale(A, x))


def test_spmv_linearity(kernel):
    A = synth.powerlaw_csr(2000, 3000, 100000, seed=31, dtype="f8", alpha=0.8)
    x1, x2 = synth.dense_vector(3000, 1, "f8"), synth.dense_vector(3000, 2, "f8")
    h = kernel.to_handle(A)
    try:
        y1, y2, y12 = kernel.mult_vec(h, x1), kernel.mult_vec(h, x2), kernel.mult_vec(h, 2.0 * x1 - x2)
    finally:
        kernel.release_handle(h)
    assert_values_close(y12, 2.0 * y1 - y2, 1e-10, _mv_scale(A, np.abs(x1) * 2 + np.abs(x2)))


@pytest.mark.parametrize("shape,nnz,dtype", [((3000, 2000), 150000, "f8"), ((500, 70000), 120000, "f4"),
                                            ((70000, 300), 200000, "f8"), ((2000, 2000), 0, "f8")])
def test_transpose_large(kernel, shape, nnz, dtype):
    A = synth.powerlaw_csr(shape[0], shape[1], nnz, seed=41, dtype=dtype, alpha=0.9)
    h = kernel.to_handle(A)
    try:
        th = kernel.transpose(h)
        T = kernel.from_handle(th)
        tth = kernel.transpose(th)
        TT = kernel.from_handle(tth)
        kernel.release_handle(th)
        kernel.release_handle(tth)
    finally:
        kernel.release_handle(h)
    R = orc.transpose(A)
    assert np.array_equal(T.rowptrs, R.rowptrs) and np.array_equal(T.colinds, R.colinds)
    assert np.array_equal(T.values, R.values)
    # (A^T)^T == A exactly (columns were sorted and unique); values widened to float64
    assert np.array_equal(TT.rowptrs, A.rowptrs) and np.array_equal(TT.colinds, A.colinds)
    assert np.array_equal(TT.values, A.values.astype(np.float64))


@pytest.mark.parametrize("bits", [8, 9])
@pytest.mark.parametrize("shape,nnz", [((700, 200), 4096 * 3), ((300, 400), 4096 * 2 + 3), ((900, 70000), 50001),
                                       ((64, 300000), 33000), ((5, 9), 17)])
def test_transpose_digit_widths_and_stability(kernel, bits, shape, nnz):
    """Both digit widths of the stable sort (1-3 passes, full and ragged last tiles) on rows with
    unsorted, duplicated columns: equal columns must keep their input order (structure.py:168-182)."""
    rng = np.random.default_rng(nnz + bits)
    nr, nc = shape
    lens = rng.multinomial(nnz, np.ones(nr) / nr)
    rp = np.zeros(nr + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    ci = rng.integers(0, nc, nnz).astype(np.int32)
    ci[rng.random(nnz) < 0.3] = nc // 3  # one hot column with many duplicates inside rows
    A = CSR(nr, nc, nnz, rp, ci, rng.standard_normal(nnz))
    kernel.set_option("radix_bits", bits)
    h = kernel.to_handle(A)
    try:
        th = kernel.transpose(h)
        T = kernel.from_handle(th)
        sh = kernel.transpose(h, False)
        S = kernel.from_handle(sh)
        kernel.release_handle(th)
        kernel.release_handle(sh)
    finally:
        kernel.release_handle(h)
        kernel.set_option("radix_bits", 0)
    R = orc.transpose(A)
    assert np.array_equal(T.rowptrs, R.rowptrs) and np.array_equal(T.colinds, R.colinds)
    assert np.array_equal(T.values, R.values)
    assert S.values is None and np.array_equal(S.rowptrs, R.rowptrs) and np.array_equal(S.colinds, R.colinds)


def test_order_columns_large_and_idempotent(kernel):
    rng = np.random.default_rng(8)
    A = synth.powerlaw_csr(3000, 50000, 200000, seed=42, dtype="f4", alpha=1.0)
    # shuffle the entries inside every row
    rows = np.repeat(np.arange(A.nrows), np.diff(A.rowptrs))
    perm = np.lexsort((rng.random(A.nnz), rows))
    S = CSR(A.nrows, A.ncols, A.nnz, A.rowptrs, A.colinds[perm], A.values[perm])
    h = kernel.to_handle(S)
    try:
        kernel.order_columns(h)
        got = kernel.from_handle(h)
        kernel.order_columns(h)
        again = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    assert np.array_equal(got.colinds, A.colinds) and np.array_equal(got.values, A.values)
    assert np.array_equal(again.colinds, got.colinds) and np.array_equal(again.values, got.values)


def _check_mm(kernel, A, B, tr, rtol):
    ref = orc.mult_abt(A, B) if tr else orc.mult_ab(A, B)
    rp, ci, vs = canonical(ref)
    ah, bh = kernel.to_handle(A), kernel.to_handle(B)
    try:
        ch = kernel.mult_abt(ah, bh) if tr else kernel.mult_ab(ah, bh)
        got = kernel.from_handle(ch)
        st = kernel.spgemm_stats(ch)
        kernel.release_handle(ch)
    finally:
        kernel.release_handle(ah)
        kernel.release_handle(bh)
    assert got.rowptrs.dtype == np.int32 and np.array_equal(got.rowptrs, rp)
    assert np.array_equal(got.colinds, ci)
    assert_values_close(got.values, vs, rtol, spgemm_terms(A, B, tr))
    assert st["out_nnz"] == ref.nnz
    return got, st


def test_spgemm_all_bins(kernel):
    """Row lengths from 0 to thousands so every symbolic bin (warp / CTA 4k / CTA 32k /
    bitmap) and every numeric bin (warp / CTA 2k / CTA 16k / dense) is exercised."""
    A = synth.powerlaw_csr(1500, 4000, 30000, seed=51, dtype="f8", alpha=1.5)
    B = synth.powerlaw_csr(4000, 30000, 150000, seed=52, dtype="f8", alpha=1.2)
    got, st = _check_mm(kernel, A, B, False, 1e-10)
    nz = np.diff(got.rowptrs)
    hist = np.histogram(nz, [0, 1, 65, 1025, 8193, 10 ** 9])[0]
    assert np.all(hist > 0), f"a numeric bin was not exercised: {hist}"
    assert st["products"] >= st["out_nnz"]


def test_spgemm_wide_global_scratch(kernel):
    "B.ncols too large for the shared-memory bitmap / dense accumulator: global scratch paths."
    A = synth.powerlaw_csr(300, 2000, 40000, seed=53, dtype="f8", alpha=1.2)
    B = synth.powerlaw_csr(2000, 2_000_000, 300000, seed=54, dtype="f8", alpha=0.5)
    _check_mm(kernel, A, B, False, 1e-10)


def _with_options(kernel, opts, fn):
    defaults = {"spgemm_esc": 1, "esc_target": 1536, "esc_budget": 0}
    for k, v in opts.items():
        kernel.set_option(k, v)
    try:
        return fn()
    finally:
        for k in opts:
            kernel.set_option(k, defaults[k])


def test_spgemm_wide_old_global_scratch(kernel):
    "the same product with the expand/sort/compress path switched off: per-CTA global scratch accumulators"
    A = synth.powerlaw_csr(300, 2000, 40000, seed=53, dtype="f8", alpha=1.2)
    B = synth.powerlaw_csr(2000, 2_000_000, 300000, seed=54, dtype="f8", alpha=0.5)
    _with_options(kernel, {"spgemm_esc": 0}, lambda: _check_mm(kernel, A, B, False, 1e-10))


@pytest.mark.parametrize("target", [16, 64, 1536, 4096])
@pytest.mark.parametrize("tr", [False, True])
def test_spgemm_esc_forced(kernel, target, tr):
    """Expand/sort/compress path forced on a product with every row size (spgemm_esc.cuh): pseudo-rows of
    `target` products -> all four reduce kernels, rows cut into up to thousands of column ranges."""
    A = synth.powerlaw_csr(1500, 4000, 30000, seed=51, dtype="f8", alpha=1.5)
    B = synth.powerlaw_csr(4000, 30000, 150000, seed=52, dtype="f8", alpha=1.2)
    if tr:
        B = B.transpose()
    got, st = _with_options(kernel, {"spgemm_esc": 2, "esc_target": target}, lambda: _check_mm(kernel, A, B, tr, 1e-10))
    assert np.diff(got.rowptrs).max() > 8192


def test_spgemm_esc_f32_products(kernel):
    "f4 x f4 products are rounded to float32 before they are summed (numba's promotion), on this path too"
    A = synth.powerlaw_csr(800, 3000, 40000, seed=61, dtype="f4", alpha=1.3)
    B = synth.powerlaw_csr(3000, 150_000, 200000, seed=62, dtype="f4", alpha=1.0)
    _with_options(kernel, {"spgemm_esc": 2}, lambda: _check_mm(kernel, A, B, False, 1e-5))


def test_spgemm_esc_hands_back_skewed_rows(kernel):
    """All of B's columns sit in the first 0.05 % of a 2M-column range: the uniform column ranges put a whole
    row into one pseudo-row; rows beyond its capacity go back to the old kernels, the others stay."""
    rng = np.random.default_rng(63)
    A = synth.powerlaw_csr(400, 2000, 40000, seed=64, dtype="f8", alpha=1.2)
    Bn = synth.powerlaw_csr(2000, 1000, 300000, seed=65, dtype="f8", alpha=0.5)
    B = CSR(2000, 2_000_000, Bn.nnz, Bn.rowptrs, Bn.colinds, Bn.values)
    got, st = _check_mm(kernel, A, B, False, 1e-10)
    lens = np.diff(B.rowptrs)
    P = np.add.reduceat(np.concatenate([lens[A.colinds], [0]]), np.minimum(A.rowptrs[:-1], A.nnz))
    P[np.diff(A.rowptrs) == 0] = 0
    assert (P > 8192).any() and ((P > 1024) & (P <= 8192)).any()


def test_spgemm_esc_hands_back_dense_rows(kernel):
    """A row with more than 256 times as many products as the result has columns (here: full rows of A against a 15 %
    dense B) is handed to the dense accumulators at once; the other rows stay on the forced expand/sort/compress path."""
    A = synth.powerlaw_csr(300, 2000, 30000, seed=72, dtype="f8", alpha=1.0)
    lens = np.diff(A.rowptrs).astype(np.int64)
    lens[7] = lens[200] = 2000                       # two full rows: P = nnz(B) = 600 000 > 256 * 2000
    rp = np.zeros(301, np.int64)
    np.cumsum(lens, out=rp[1:])
    rng = np.random.default_rng(73)
    cols = synth.stratified_columns(lens, 2000, rng)
    A = CSR(300, 2000, int(rp[-1]), rp, cols, rng.normal(size=int(rp[-1])))
    B = synth.powerlaw_csr(2000, 2000, 600_000, seed=74, dtype="f8", alpha=0.3)
    got, st = _with_options(kernel, {"spgemm_esc": 2}, lambda: _check_mm(kernel, A, B, False, 1e-10))
    nz = np.diff(got.rowptrs)
    assert nz[7] == 2000 or nz[7] > 1900


def test_spgemm_esc_popular_column(kernel):
    """Every row of B holds column 7: an output element collects a product from EVERY entry of A's row (15 000 in one
    pseudo-row, beyond the 8192 the sort/merge kernels take); the ranges are narrow, so the hash kernel with warp
    pre-aggregation takes such pseudo-rows (k_esc_check, k_esc_reduce<..., AGG>)."""
    rng = np.random.default_rng(77)
    lens = np.full(50, 40, np.int64)
    lens[5], lens[31] = 15000, 9000
    rp = np.zeros(51, np.int64)
    np.cumsum(lens, out=rp[1:])
    A = CSR(50, 20000, int(rp[-1]), rp, synth.stratified_columns(lens, 20000, rng), rng.normal(size=int(rp[-1])))
    bl = np.full(20000, 4, np.int64)
    brp = np.zeros(20001, np.int64)
    np.cumsum(bl, out=brp[1:])
    bc = np.sort(np.concatenate([np.full((20000, 1), 7), rng.integers(8, 200_000, (20000, 3))], axis=1), axis=1).astype(np.int32)
    B = CSR(20000, 200_000, 80000, brp, bc.reshape(-1), rng.normal(size=80000))
    for target in (16, 1536):
        _with_options(kernel, {"spgemm_esc": 2, "esc_target": target}, lambda: _check_mm(kernel, A, B, False, 1e-10))


def test_spgemm_long_rows_of_a(kernel):
    "rows of A beyond 65 536 entries: their product counts are summed piecewise (k_row_products_long)"
    lens = np.full(40, 50, np.int64)
    lens[3], lens[21] = 150_000, 70_001
    rp = np.zeros(41, np.int64)
    np.cumsum(lens, out=rp[1:])
    rng = np.random.default_rng(75)
    cols = synth.stratified_columns(lens, 200_000, rng)
    A = CSR(40, 200_000, int(rp[-1]), rp, cols, rng.normal(size=int(rp[-1])))
    B = synth.powerlaw_csr(200_000, 3000, 600_000, seed=76, dtype="f8", alpha=0.5)
    got, st = _check_mm(kernel, A, B, False, 1e-10)
    lb = np.diff(B.rowptrs).astype(np.int64)
    assert st["products"] == int(lb[A.colinds].sum())


def test_spgemm_esc_declines_over_budget(kernel):
    A = synth.powerlaw_csr(300, 2000, 40000, seed=53, dtype="f8", alpha=1.2)
    B = synth.powerlaw_csr(2000, 2_000_000, 300000, seed=54, dtype="f8", alpha=0.5)
    _with_options(kernel, {"esc_budget": 4096}, lambda: _check_mm(kernel, A, B, False, 1e-10))


def test_spgemm_esc_columns_pile_up(kernel):
    """B uses 40 distinct columns of 300 000: every pseudo-row has buckets of hundreds of equal columns, the
    bucket-sort kernel flags it and the hash kernel takes over (k_esc_reduce)."""
    rng = np.random.default_rng(68)
    A = synth.powerlaw_csr(300, 800, 30000, seed=69, dtype="f8", alpha=1.0)
    hot = np.sort(rng.choice(300_000, 40, replace=False)).astype(np.int32)
    lens = rng.integers(1, 30, 800)
    rp = np.zeros(801, np.int64)
    np.cumsum(lens, out=rp[1:])
    cols = np.concatenate([np.sort(rng.choice(hot, int(l), replace=False)) for l in lens]).astype(np.int32)
    B = CSR(800, 300_000, int(rp[-1]), rp, cols, rng.normal(size=int(rp[-1])))
    for target in (64, 1536):
        _with_options(kernel, {"spgemm_esc": 2, "esc_target": target}, lambda: _check_mm(kernel, A, B, False, 1e-10))


def test_spgemm_esc_unsorted_duplicate_columns(kernel):
    "B rows with unsorted and repeated columns (legal CSR for the reference: multiply.py:60-129 has no order assumption)"
    rng = np.random.default_rng(66)
    A = synth.powerlaw_csr(200, 500, 20000, seed=67, dtype="f8", alpha=1.0)
    lens = rng.integers(0, 120, 500)
    rp = np.zeros(501, np.int64)
    np.cumsum(lens, out=rp[1:])
    cols = rng.integers(0, 300_000, int(rp[-1])).astype(np.int32)
    idx = np.arange(0, len(cols) - 1, 7)
    cols[idx] = cols[idx + 1]      # repeated columns, mostly inside one row
    B = CSR(500, 300_000, int(rp[-1]), rp, cols, rng.normal(size=int(rp[-1])))
    _with_options(kernel, {"spgemm_esc": 2, "esc_target": 256}, lambda: _check_mm(kernel, A, B, False, 1e-10))


@pytest.mark.parametrize("dtype", ["f8", "f4"])
def test_item_item_abt(kernel, dtype):
    "configs[2] shape, scaled: M = ratings^T, M M^T"
    R = synth.cfg3_ratings(0.02)
    if dtype == "f4":
        R = CSR(R.nrows, R.ncols, R.nnz, R.rowptrs, R.colinds, R.values.astype(np.float32))
    M = R.transpose()          # float64 values either way (structure.py:177)
    assert M.values.dtype == np.float64
    _check_mm(kernel, M, M, True, 1e-10)


@pytest.mark.parametrize("chunk_prod", [-1, 0, 60000])
def test_item_item_owner_path(kernel, chunk_prod):
    """Rows with > 8192 output entries and a sorted (transposed) right operand take the
    owner-computes dense kernel: no atomics.  Unchunked (own_chunk_prod < 0; the default rule only
    chunks when one row would be the critical path) every output element is summed in the
    reference's own order, so the VALUES are bit-identical to the oracle too.  With heavy rows cut into chunks of A entries (symbolic and
    numeric pass, forced here with 60000 products per chunk) the structure stays exact and the
    values differ only by the order in which the chunk sums are added."""
    R = synth.powerlaw_csr(16000, 12000, 1_600_000, seed=81, dtype="f8", alpha=0.5, cap=600, min_len=20, col_skew=2.0)
    M = R.transpose()
    ref = orc.mult_abt(M, M)
    rp, ci, vs = canonical(ref)
    assert np.diff(rp).max() > 8192
    kernel.set_option("own_chunk_prod", chunk_prod)
    kernel.set_option("spgemm_fixed", 0)   # positive ratings would otherwise take the fixed-point kernel
    mh = kernel.to_handle(M)
    try:
        ch = kernel.mult_abt(mh, mh)
        got = kernel.from_handle(ch)
        st = kernel.spgemm_stats(ch)
        kernel.release_handle(ch)
    finally:
        kernel.release_handle(mh)
        kernel.set_option("own_chunk_prod", 0)
        kernel.set_option("spgemm_fixed", 1)
    assert np.array_equal(got.rowptrs, rp) and np.array_equal(got.colinds, ci)
    assert st["out_nnz"] == ref.nnz and st["dense_path"] == "owner"
    heavy = np.repeat(np.diff(rp) > 8192, np.diff(rp))
    if chunk_prod < 0:
        assert np.array_equal(got.values[heavy], vs[heavy]), "owner path must reproduce the reference bit for bit"
    elif chunk_prod > 0:
        assert not np.array_equal(got.values[heavy], vs[heavy]), "chunking was requested but did not happen"
    assert_values_close(got.values, vs, 1e-10)   # positive terms: the bound is the value itself


def _abt(kernel, M, **opts):
    for k, v in opts.items():
        kernel.set_option(k, v)
    mh = kernel.to_handle(M)
    try:
        ch = kernel.mult_abt(mh, mh)
        got = kernel.from_handle(ch)
        st = kernel.spgemm_stats(ch)
        kernel.release_handle(ch)
    finally:
        kernel.release_handle(mh)
        kernel.set_option("own_chunk_prod", 0)
        kernel.set_option("spgemm_fixed", 1)
    return got, st


@pytest.mark.parametrize("dtype", ["f8", "f4"])
def test_item_item_fixed_point_path(kernel, dtype):
    """Non-negative values of bounded range (ratings): heavy rows are accumulated in 64-bit fixed point with
    native shared-memory atomics.  Structure exact; EVERY value within rtol 1e-10 (f64) of the oracle with
    no absolute slack; and, integer sums being order-independent, cutting rows into chunks changes nothing."""
    R = synth.powerlaw_csr(16000, 12000, 1_600_000, seed=81, dtype=dtype, alpha=0.5, cap=600, min_len=20, col_skew=2.0)
    M = R.transpose()            # values become float64 (of float32 ratings for "f4")
    ref = orc.mult_abt(M, M)
    rp, ci, vs = canonical(ref)
    got, st = _abt(kernel, M)
    assert st["dense_path"] == "fixed"
    assert np.array_equal(got.rowptrs, rp) and np.array_equal(got.colinds, ci)
    assert np.allclose(got.values, vs, rtol=1e-10, atol=0.0)
    chunked, st2 = _abt(kernel, M, own_chunk_prod=60000)
    assert st2["dense_path"] == "fixed"
    heavy = np.repeat(np.diff(rp) > 8192, np.diff(rp))   # (lighter rows use float64 hash accumulators: order varies)
    assert np.array_equal(chunked.colinds, got.colinds) and np.array_equal(chunked.values[heavy], got.values[heavy])
    assert np.allclose(chunked.values, vs, rtol=1e-10, atol=0.0)


def test_fixed_point_signed_values(kernel):
    """Mixed signs with a bounded range of magnitudes take the fixed-point kernel (two's-complement words): every
    element within 1e-10 of the magnitude that was summed, sum_k |a_ik||b_kj| -- the bound of a re-ordered sum --
    and the result does not depend on the chunking."""
    R = synth.powerlaw_csr(16000, 12000, 1_600_000, seed=81, dtype="f8", alpha=0.5, cap=600, min_len=20, col_skew=2.0)
    v = R.values.copy()
    v[::7] *= -1.0
    v[1::3] *= -1.0
    M = CSR(R.nrows, R.ncols, R.nnz, R.rowptrs, R.colinds, v).transpose()
    ref = orc.mult_abt(M, M)
    rp, ci, vs = canonical(ref)
    terms = spgemm_terms(M, M, transpose=True)
    got, st = _abt(kernel, M)
    assert st["dense_path"] == "fixed"
    assert np.array_equal(got.rowptrs, rp) and np.array_equal(got.colinds, ci)
    assert_values_close(got.values, vs, 1e-10, terms)
    chunked, st2 = _abt(kernel, M, own_chunk_prod=30000)
    heavy = np.repeat(np.diff(rp) > 8192, np.diff(rp))      # the rows of the dense (fixed-point) path
    assert st2["dense_path"] == "fixed" and np.array_equal(chunked.values[heavy], got.values[heavy])


def _ratings_variant(case):
    R = synth.powerlaw_csr(16000, 12000, 1_600_000, seed=81, dtype="f8", alpha=0.5, cap=600, min_len=20, col_skew=2.0)
    v = R.values.copy()
    lens = np.diff(R.rowptrs)
    rows = R.rowptrs[:-1].astype(np.int64)
    if case == "centred":          # normalize_rows('center'): values arbitrarily close to 0, mixed signs
        v = v - np.repeat(np.add.reduceat(v, rows) / np.maximum(lens, 1), lens)
    elif case == "unit":           # normalize_rows('unit'): magnitudes differ by the rows' norms
        v = v / np.repeat(np.sqrt(np.add.reduceat(v * v, rows)), lens)
    elif case == "wide_range":
        v[::200] *= 1e-7
    elif case == "very_wide_range":
        v[::5] *= 1e-7
    elif case == "nonfinite":
        v[12345] = np.inf
    return CSR(R.nrows, R.ncols, R.nnz, R.rowptrs, R.colinds, v).transpose()


@pytest.mark.parametrize("case", ["centred", "unit", "wide_range"])
def test_fixed_point_any_value_range(kernel, case):
    """Round 2: the operands are equilibrated by exact powers of two (A by row, B by column) and products below
    the accumulator's grid go through an exact side list, so mean-centred and unit-normalised ratings -- the
    inputs of item-item similarity -- and any other finite values take the fixed-point kernel, with every element
    within 1e-10 of sum_k |a_ik||b_kj| (no absolute term)."""
    M = _ratings_variant(case)
    ref = orc.mult_abt(M, M)
    rp, ci, vs = canonical(ref)
    terms = spgemm_terms(M, M, transpose=True)
    got, st = _abt(kernel, M)
    assert st["dense_path"] == "fixed"
    assert np.array_equal(got.rowptrs, rp) and np.array_equal(got.colinds, ci)
    assert_values_close(got.values, vs, 1e-10, terms)
    if case == "wide_range":
        assert st["side_list"] > 0
    if case == "unit":
        assert st["side_list"] == 0   # equilibration alone brings the products within the accumulator's range
    chunked, st2 = _abt(kernel, M, own_chunk_prod=30000)
    assert st2["dense_path"] == "fixed"
    assert_values_close(chunked.values, vs, 1e-10, terms)


@pytest.mark.parametrize("case", ["nonfinite", "side_list_overflow", "very_wide_range"])
def test_fixed_point_gate_falls_back_to_owner(kernel, case):
    """A non-finite value, or more coarse products than the side list holds, send the heavy rows to the
    owner-computes kernel (bit-identical to the oracle)."""
    # (a 64-entry list overflows at once; "very_wide_range" -- a third of the products below the grid -- overflows
    # the default one: 1/32 of the products)
    M = _ratings_variant({"side_list_overflow": "wide_range"}.get(case, case))
    ref = orc.mult_abt(M, M)
    rp, ci, vs = canonical(ref)
    try:
        got, st = _abt(kernel, M, own_chunk_prod=-1, **({"fix_tiny_cap": 64} if case == "side_list_overflow" else {}))
    finally:
        kernel.set_option("fix_tiny_cap", 0)
    assert st["dense_path"] == "owner"
    assert np.array_equal(got.rowptrs, rp) and np.array_equal(got.colinds, ci)
    heavy = np.repeat(np.diff(rp) > 8192, np.diff(rp))
    assert np.array_equal(got.values[heavy], vs[heavy], equal_nan=True)


def test_virtual_ranks_row_blocks(kernel):
    """N ranks emulated as N row shards on one GPU; the concatenation equals
    the unsharded result (structure exactly; values up to atomic ordering)."""
    A = synth.powerlaw_csr(3000, 2500, 90000, seed=61, dtype="f8", alpha=1.0)
    B = synth.powerlaw_csr(2500, 3500, 80000, seed=62, dtype="f8", alpha=0.8)
    x = synth.dense_vector(2500, 63, "f8")
    full = A.multiply(B)
    yfull = A.mult_vec(x)
    for n in (2, 4, 8):
        cuts = partition_rows(A.rowptrs, n)
        shards = [A.subset_rows(cuts[i], cuts[i + 1]) for i in range(n)]
        parts = [s.multiply(B) for s in shards]
        C = CSR._assemble_shards(parts)
        assert np.array_equal(C.rowptrs, full.rowptrs) and np.array_equal(C.colinds, full.colinds)
        assert_values_close(C.values, full.values, 1e-12)   # positive terms
        y = np.concatenate([s.mult_vec(x) for s in shards])
        assert_values_close(y, yfull, 1e-12, _mv_scale(A, x))
        # the same through device-side row slices of one resident handle
        h = kernel.to_handle(A)
        try:
            ys = []
            for i in range(n):
                sh = kernel.subset_rows(h, cuts[i], cuts[i + 1])
                ys.append(kernel.mult_vec(sh, x))
                kernel.release_handle(sh)
        finally:
            kernel.release_handle(h)
        assert_values_close(np.concatenate(ys), yfull, 1e-12, _mv_scale(A, x))


def test_structure_only_product(kernel):
    "Superset of the numba kernel (which cannot type value-less inputs): values count as 1, like the MKL kernel."
    A = synth.powerlaw_csr(200, 300, 4000, seed=71, dtype="f8", values=False)
    B = synth.powerlaw_csr(300, 250, 5000, seed=72, dtype="f8", values=False)
    ones = lambda m: CSR(m.nrows, m.ncols, m.nnz, m.rowptrs, m.colinds, np.ones(m.nnz))
    ref = canonical(orc.mult_ab(ones(A), ones(B)))
    ah, bh = kernel.to_handle(A), kernel.to_handle(B)
    ch = kernel.mult_ab(ah, bh)
    got = kernel.from_handle(ch)
    for hh in (ah, bh, ch):
        kernel.release_handle(hh)
    assert np.array_equal(got.rowptrs, ref[0]) and np.array_equal(got.colinds, ref[1])
    assert np.array_equal(got.values, ref[2])     # small integer counts: exact


def test_keep_resident_handle_cache(kernel):
    "CSR-level device residency: one upload for many mult_vec calls; dropped on new values / GC."
    from csr_b200 import csr as csr_mod
    A = synth.powerlaw_csr(500, 400, 6000, seed=91, dtype="f8").keep_resident()
    x = synth.dense_vector(400, 92, "f8")
    y1 = A.mult_vec(x)
    h1 = csr_mod._resident._live[id(A)][1]
    y2 = A.mult_vec(x)
    assert csr_mod._resident._live[id(A)][1] is h1 and h1.H
    assert np.array_equal(y1, y2)
    assert_values_close(y1, orc.mult_vec(A, x), 1e-10, _mv_scale(A, x))
    A.values = A.values * 2.0            # re-assigned values invalidate the cached handle
    assert not h1.H and id(A) not in csr_mod._resident._live
    assert_values_close(A.mult_vec(x), 2.0 * y1, 1e-12, _mv_scale(A, x))
    h2 = csr_mod._resident._live[id(A)][1]
    key = id(A)
    del A
    import gc
    gc.collect()
    assert key not in csr_mod._resident._live and not h2.H


def test_large_export_goes_through_the_threaded_copy(kernel):
    """Arrays above 128 MB leave the device through pinned slots filled and drained by several host threads
    (context.cu copy_to_host): the round trip must be bit-exact, with a tail chunk that is not slot-sized."""
    n = 36_000_001
    rng = np.random.default_rng(91)
    rp = np.zeros(1001, np.int64)
    rp[1:] = np.sort(rng.integers(0, n, 1000))
    rp[-1] = n
    cols = rng.integers(0, 50_000, n).astype(np.int32)
    vals = rng.standard_normal(n)
    A = CSR(1000, 50_000, n, rp, cols, vals)
    h = kernel.to_handle(A)
    try:
        B = kernel.from_handle(h)
    finally:
        kernel.release_handle(h)
    assert np.array_equal(B.rowptrs, A.rowptrs) and np.array_equal(B.colinds, cols) and np.array_equal(B.values, vals)


def test_released_handle_is_rejected(kernel):
    h = kernel.to_handle(CSR.empty(3, 3))
    kernel.release_handle(h)
    kernel.release_handle(h)           # second release is a no-op, like mkl_h
    with pytest.raises(ValueError):
        kernel.mult_vec(h, np.zeros(3))


def test_shape_errors(kernel):
    a = kernel.to_handle(CSR.empty(3, 4))
    b = kernel.to_handle(CSR.empty(5, 6))
    try:
        with pytest.raises(AssertionError):
            kernel.mult_ab(a, b)         # multiply.py:26
        with pytest.raises(AssertionError):
            kernel.mult_abt(a, b)        # multiply.py:53
        with pytest.raises(ValueError):
            kernel.mult_vec(a, np.zeros(3))
    finally:
        kernel.release_handle(a)
        kernel.release_handle(b)
