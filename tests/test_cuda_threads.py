"""
GPU: the library is re-entrant (include/csrk.h "Threading"): the reference's kernels are ``nogil``
(csr/kernels/numba/__init__.py:55, multiply.py:13,41), so several Python threads may call them at once on shared,
read-only handles, and ctypes releases the GIL around every call into libcsr_cuda.so.  Threads here hammer
``mult_vec`` (both SpMV kernels, including the first call that builds a plan), ``mult_ab`` / ``mult_abt``,
``transpose`` and handle creation/release concurrently; every result must equal the one a single thread gets.
"""

import threading

import numpy as np
import pytest

from csr_b200 import synth
from oracle import oracle as orc
from util import canonical, assert_values_close, spmv_terms

pytestmark = pytest.mark.gpu

NTHREADS = 4
ROUNDS = 6


def _run_threads(fn):
    errs = []

    def wrap(i):
        try:
            fn(i)
        except BaseException as e:   # noqa: BLE001 - reported by the main thread
            errs.append((i, repr(e)))
    ts = [threading.Thread(target=wrap, args=(i,)) for i in range(NTHREADS)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


@pytest.mark.parametrize("mode", [1, 2], ids=["tile-kernel", "slab-kernel"])
def test_concurrent_mult_vec_on_a_shared_handle(kernel, mode):
    A = synth.powerlaw_csr(20000, 30000, 1_500_000, seed=91, dtype="f4", alpha=1.0)
    xs = [synth.dense_vector(A.ncols, 100 + i, "f4") for i in range(NTHREADS)]
    refs = [orc.mult_vec(A, x) for x in xs]
    terms = [spmv_terms(A, x) for x in xs]
    kernel.set_option("spmv_mode", mode)
    h = kernel.to_handle(A)          # no plan yet: the first calls race to build it
    try:
        def work(i):
            first = None
            for _ in range(ROUNDS):
                y = kernel.mult_vec(h, xs[i])
                assert_values_close(y, refs[i], 1e-5, terms[i])
                if first is None:
                    first = y
                assert np.array_equal(y, first), "mult_vec must be deterministic under concurrency"
        _run_threads(work)
        info = kernel.spmv_plan_info(h, 4)
        assert info["kernel"] == ("stream" if mode == 2 else "tile")
    finally:
        kernel.release_handle(h)
        kernel.set_option("spmv_mode", 0)


def test_concurrent_products_transposes_and_handle_churn(kernel):
    A = synth.powerlaw_csr(3000, 2500, 90000, seed=92, dtype="f8", alpha=1.0)
    B = synth.powerlaw_csr(2500, 3500, 80000, seed=93, dtype="f8", alpha=0.8)
    ab = canonical(orc.mult_ab(A, B))
    aat = canonical(orc.mult_abt(A, A))
    at = orc.transpose(A)
    ah, bh = kernel.to_handle(A), kernel.to_handle(B)
    try:
        def work(i):
            for r in range(ROUNDS):
                which = (i + r) % 4
                if which == 0:
                    ch = kernel.mult_ab(ah, bh)
                    C = kernel.from_handle(ch)
                    kernel.release_handle(ch)
                    assert np.array_equal(C.rowptrs, ab[0]) and np.array_equal(C.colinds, ab[1])
                    assert_values_close(C.values, ab[2], 1e-10)
                elif which == 1:
                    ch = kernel.mult_abt(ah, ah)
                    C = kernel.from_handle(ch)
                    kernel.release_handle(ch)
                    assert np.array_equal(C.rowptrs, aat[0]) and np.array_equal(C.colinds, aat[1])
                    assert_values_close(C.values, aat[2], 1e-10)
                elif which == 2:
                    th = kernel.transpose(ah)
                    T = kernel.from_handle(th)
                    kernel.release_handle(th)
                    assert np.array_equal(T.rowptrs, at.rowptrs) and np.array_equal(T.colinds, at.colinds)
                    assert np.array_equal(T.values, at.values)
                else:
                    h2 = kernel.to_handle(B)      # create / export / release while the others compute
                    B2 = kernel.from_handle(h2)
                    kernel.release_handle(h2)
                    assert np.array_equal(B2.colinds, B.colinds) and np.array_equal(B2.values, B.values)
        _run_threads(work)
    finally:
        kernel.release_handle(ah)
        kernel.release_handle(bh)
