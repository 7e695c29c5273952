"""
CPU: a numpy model of the SpGEMM fixed-point accumulator (csrc/spgemm.cu, k_num_fixed and the gate in
spgemm_run), checking the error bound the gate promises.

Model: a row with `length` stored entries gets hb = ceil(log2(length + 1)) bits of headroom; with
2^x >= max|a*b| (frexp) every product is rounded once to a multiple of 2^-(62 - x - hb) and the integers
are added exactly (two 32-bit words with a carry on the device, Python ints here).  The gate admits
non-negative finite values with max|a*b| / min|a*b| <= 2^(26 - hb); the claim is a relative error of at
most 2^-35 for EVERY output element, whatever the number of terms (up to `length`) and their order.
"""

import math

import numpy as np
import pytest


def fixed_point_sum(terms, length, pmax):
    hb = int(length).bit_length()                 # ceil(log2(length + 1)), as headroom_bits()
    _, x = math.frexp(pmax)                       # pmax <= 2^x
    sh = 62 - x - hb
    scaled = np.rint(np.ldexp(terms, sh))         # __double2ll_rn(p * 2^sh): exact scaling, one rounding
    assert np.all(np.abs(scaled) < 2.0 ** 62)
    total = sum(int(v) for v in scaled)           # exact integer addition (order-independent)
    assert abs(total) < 2 ** 63                   # fits the 64-bit accumulator
    return math.ldexp(float(total), -sh), hb      # (double)v * 2^-sh


@pytest.mark.parametrize("length", [9000, 100000, 1 << 20])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_error_bound_at_the_gate_limit(length, seed):
    rng = np.random.default_rng(seed)
    hb = int(length).bit_length()
    ratio = 2.0 ** (26 - hb)                      # the widest value range the gate admits for this row length
    pmax = 10.0 ** rng.uniform(-30, 30)
    pmin = pmax / ratio
    for nterms in (1, 2, 17, length):             # one output element may collect up to `length` products
        # worst case for a relative bound: all terms at the small end, plus the extremes
        for terms in (np.full(nterms, pmin), rng.uniform(pmin, pmax, nterms),
                      np.concatenate([[pmax], np.full(nterms - 1, pmin)])):
            got, _ = fixed_point_sum(terms, length, pmax)
            exact = math.fsum(terms)
            assert abs(got - exact) <= 2.0 ** -35 * exact, (length, nterms, got, exact)


def test_order_independence_and_chunking():
    rng = np.random.default_rng(7)
    terms = rng.uniform(0.25, 25.0, 50000)        # products of ratings in [0.5, 5]
    a, _ = fixed_point_sum(terms, 60000, 25.0)
    b, _ = fixed_point_sum(rng.permutation(terms), 60000, 25.0)
    # chunks: integer partial sums added afterwards give the same integer
    hb = (60000).bit_length()
    sh = 62 - math.frexp(25.0)[1] - hb
    parts = [sum(int(v) for v in np.rint(np.ldexp(c, sh))) for c in np.array_split(terms, 7)]
    c = math.ldexp(float(sum(parts)), -sh)
    assert a == b == c


def test_gate_rejects_what_the_bound_cannot_cover():
    # one decade more dynamic range than admitted: the bound is no longer guaranteed (and is in fact broken)
    length = 100000
    hb = int(length).bit_length()
    pmax = 1.0
    pmin = pmax / (2.0 ** (26 - hb) * 2.0 ** 12)
    got, _ = fixed_point_sum(np.array([pmin * 1.37]), length, pmax)
    assert abs(got - pmin * 1.37) > 2.0 ** -35 * pmin * 1.37
