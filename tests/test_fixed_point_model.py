"""
CPU: a numpy model of the SpGEMM fixed-point accumulator (csrc/spgemm.cu: k_num_fixed, k_fix_tiny and the
equilibration in spgemm_run), checking the error bound it promises.

Model (round 2): the operands are equilibrated by exact powers of two -- a_ik * 2^-ea_i with
max_k |a_ik| < 2^ea_i, b_kj * 2^-eb_j with max_k |b_kj| < 2^eb_j -- so every product is below 1 in magnitude.
A row with `length` stored entries gets hb = ceil(log2(length + 1)) bits of headroom and accumulates
T = rint(p' * 2^(62 - hb)) as exact integers (two 32-bit words with a carry on the device, Python ints
here); a product with |T| < 2^35 is NOT accumulated but added exactly, in float64, to the finished element
(the side list).  The claim: EVERY output element is within 2^-35 of sum |terms|, whatever the values, the
number of terms (up to `length`) and their order.
"""

import math

import numpy as np
import pytest


def fixed_point_sum(terms, length):
    """terms: equilibrated products (|p'| < 1) of ONE output element of a row with `length` entries."""
    hb = int(length).bit_length()                 # ceil(log2(length + 1)), as headroom_bits()
    sh = 62 - hb
    scaled = np.rint(np.ldexp(terms, sh))         # __double2ll_rn(p' * 2^sh): exact scaling, one rounding
    big = np.abs(scaled) >= 2.0 ** 35
    total = sum(int(v) for v in scaled[big])      # exact integer addition (order-independent)
    assert abs(total) < 2 ** 63                   # fits the 64-bit accumulator
    out = math.ldexp(float(total), -sh)           # (double)v * 2^-sh
    for t in terms[~big]:                         # k_fix_tiny: float64 atomicAdd, one at a time
        out += float(t)
    return out, int((~big).sum())


@pytest.mark.parametrize("length", [9000, 100000, 1 << 20])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_error_bound_for_any_value_range(length, seed):
    rng = np.random.default_rng(seed)
    for nterms in (1, 2, 17, 4096):               # one output element may collect up to `length` products
        for terms in (np.full(nterms, 1.0 - 2.0 ** -20),                         # all at the top: needs the headroom
                      rng.uniform(-1, 1, nterms) * 10.0 ** rng.uniform(-12, 0, nterms),   # twelve decades, mixed signs
                      np.full(nterms, 1.37e-9),                                   # all below the grid: side list only
                      np.concatenate([[0.9], np.full(nterms - 1, -3.1e-11)])):
            got, _ = fixed_point_sum(np.asarray(terms, np.float64), length)
            exact = math.fsum(terms)
            assert abs(got - exact) <= 2.0 ** -35 * math.fsum(np.abs(terms)), (length, nterms, got, exact)


def test_headroom_is_enough_for_a_full_row():
    length = 100000
    got, small = fixed_point_sum(np.full(length, 1.0 - 2.0 ** -30), length)
    assert small == 0 and abs(got - length * (1.0 - 2.0 ** -30)) <= 2.0 ** -35 * length


def test_order_independence_and_chunking():
    rng = np.random.default_rng(7)
    terms = rng.uniform(0.01, 1.0, 50000)         # equilibrated products of ratings
    a, _ = fixed_point_sum(terms, 60000)
    b, _ = fixed_point_sum(rng.permutation(terms), 60000)
    # chunks: integer partial sums added afterwards give the same integer
    sh = 62 - (60000).bit_length()
    parts = [sum(int(v) for v in np.rint(np.ldexp(c, sh))) for c in np.array_split(terms, 7)]
    c = math.ldexp(float(sum(parts)), -sh)
    assert a == b == c


def test_the_side_list_is_what_keeps_small_terms():
    # without it a product below the grid would simply vanish: the bound would be broken
    length = 100000
    sh = 62 - int(length).bit_length()
    p = 1.37 * 2.0 ** -(sh + 2)
    assert np.rint(np.ldexp(p, sh)) == 0.0
    got, small = fixed_point_sum(np.array([p]), length)
    assert small == 1 and got == p
