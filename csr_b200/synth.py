"""
Seeded synthetic workloads for BASELINE.json's configs.

All generators are NumPy ``default_rng(seed)`` and vectorised (a 100M-nnz matrix
takes seconds), produce unique, ascending columns inside every row, values
``uniform(0.5, 5.0)`` (rating-like, never zero) and ``x ~ N(0, 1)``.

Row lengths follow a rank-size power law ``len(rank r) ~ r**-alpha`` scaled to the
requested nnz, clipped to ``[min_len, cap]`` and assigned to rows in random order.
Columns of a row of length L are a stratified sample of the column space: the
k-th entry falls in ``[k*C/L, (k+1)*C/L)`` (unique and sorted by construction);
``col_skew`` = g > 1 warps the strata by ``t -> t**g`` so that low column ids are
popular (item-popularity skew), followed by a strictly-increasing fix-up.
"""

from __future__ import annotations

import numpy as np

from .csr import CSR


def powerlaw_lengths(nrows: int, nnz: int, alpha: float, cap: int, min_len: int, rng) -> np.ndarray:
    """Integer row lengths with sum exactly ``nnz`` (int64[nrows])."""
    assert nrows > 0 and 0 <= min_len <= cap
    assert nrows * min_len <= nnz <= nrows * cap, "nnz outside [nrows*min_len, nrows*cap]"
    w = np.arange(1, nrows + 1, dtype=np.float64) ** (-alpha)

    def total(c):
        return np.clip(np.floor(c * w), min_len, cap).sum()

    lo, hi = 0.0, 1.0
    while total(hi) < nnz:
        hi *= 2.0
    for _ in range(100):
        mid = 0.5 * (lo + hi)
        if total(mid) < nnz:
            lo = mid
        else:
            hi = mid
    lens = np.clip(np.floor(lo * w), min_len, cap).astype(np.int64)
    rem = int(nnz - lens.sum())
    # hand the remainder out one entry at a time to rows that still have room
    while rem > 0:
        room = np.flatnonzero(lens < cap)
        take = room[:rem]
        lens[take] += 1
        rem -= len(take)
    return rng.permutation(lens)


def stratified_columns(lens: np.ndarray, ncols: int, rng, col_skew: float = 1.0,
                       chunk_nnz: int = 1 << 24) -> np.ndarray:
    """Unique ascending int32 columns for every row (rows with len <= ncols)."""
    nnz = int(lens.sum())
    out = np.empty(nnz, np.int32)
    rp = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    r0 = 0
    nrows = len(lens)
    while r0 < nrows:
        r1 = int(np.searchsorted(rp, rp[r0] + chunk_nnz, side="right"))
        r1 = min(max(r1 - 1, r0 + 1), nrows)
        ln = lens[r0:r1]
        n = int(rp[r1] - rp[r0])
        if n:
            L = np.repeat(ln, ln)
            k = np.arange(n, dtype=np.int64) - np.repeat(rp[r0:r1] - rp[r0], ln)
            u = rng.random(n)
            if col_skew == 1.0:
                lo = (k * ncols) // L
                hi = ((k + 1) * ncols) // L
                c = lo + (u * (hi - lo)).astype(np.int64)
            else:
                t = (k + u) / L
                c = np.minimum((ncols * t ** col_skew).astype(np.int64), ncols - 1)
                # strictly increasing inside each row: c_k = max_{j<=k}(c_j - j) + k
                big = 4 * (ncols + int(ln.max()))
                row = np.repeat(np.arange(r1 - r0, dtype=np.int64), ln)
                v = c - k + row * big
                np.maximum.accumulate(v, out=v)
                c = v - row * big + k
            out[rp[r0]:rp[r1]] = c
        r0 = r1
    return out


def powerlaw_csr(nrows: int, ncols: int, nnz: int, *, seed: int, dtype="f4", alpha: float = 1.0,
                 cap: int | None = None, min_len: int = 0, col_skew: float = 1.0, values: bool = True) -> CSR:
    rng = np.random.default_rng(seed)
    cap = ncols if cap is None else min(cap, ncols)
    lens = powerlaw_lengths(nrows, nnz, alpha, cap, min_len, rng)
    cols = stratified_columns(lens, ncols, rng, col_skew)
    rps = np.zeros(nrows + 1, np.int64)
    np.cumsum(lens, out=rps[1:])
    vals = None
    if values:
        vals = rng.random(nnz, dtype=np.float32 if np.dtype(dtype) == np.float32 else np.float64)
        vals *= 4.5
        vals += 0.5
        vals = vals.astype(dtype, copy=False)
    return CSR(nrows, ncols, nnz, rps, cols, vals)


def dense_vector(n: int, seed: int, dtype="f4") -> np.ndarray:
    return np.random.default_rng(seed).standard_normal(n).astype(dtype)


# ---- BASELINE.json configs ---------------------------------------------------
# scale < 1 shrinks rows, columns and nnz together (parity tests use small scales)

def cfg1_movielens(scale: float = 1.0) -> CSR:
    "6040 x 3706, 1,000,209 nnz, float64: ML-1M-shaped (min 20 ratings per user)."
    nr, nc = max(int(6040 * scale), 8), max(int(3706 * scale), 8)
    nnz = min(int(1000209 * scale * scale), nr * nc // 2)
    return powerlaw_csr(nr, nc, nnz, seed=1, dtype="f8", alpha=0.7, cap=int(nc * 0.63),
                        min_len=min(20, nnz // nr), col_skew=2.0)


def cfg2_spmv(scale: float = 1.0, col_skew: float = 1.0, seed: int = 2, dtype="f4") -> CSR:
    "1M x 1M, 100M nnz, float32, power-law rows (alpha 1, mean 100)."
    n = max(int(1_000_000 * scale), 64)
    return powerlaw_csr(n, n, 100 * n, seed=seed, dtype=dtype, alpha=1.0, col_skew=col_skew)


def cfg3_ratings(scale: float = 1.0) -> CSR:
    """100k users x 50k items, 20M nnz, float64.  Item-item similarity is
    ``M.multiply(M, transpose=True)`` with ``M = ratings.transpose()``.  User lengths
    are capped and item popularity is skewed (t**3 strata) so that out-nnz stays below 2**31
    (measured: Z = 1.6e9, 65 % dense, P = 7.3e9 products at full scale)."""
    nu, ni = max(int(100_000 * scale), 16), max(int(50_000 * scale), 16)
    nnz = min(200 * nu, nu * ni // 4)
    mean = max(nnz // nu, 1)
    return powerlaw_csr(nu, ni, nnz, seed=3, dtype="f8", alpha=0.5,
                        cap=max(ni // 25, min(ni, 4 * mean)), min_len=min(20, mean // 2), col_skew=3.0)


def cfg4_square(scale: float = 1.0, dtype="f8") -> CSR:
    "5M x 5M, 500M nnz, near-uniform columns (general mult_ab + transpose)."
    n = max(int(5_000_000 * scale), 64)
    return powerlaw_csr(n, n, 100 * n, seed=4, dtype=dtype, alpha=0.8, cap=max(n // 50, 8))
