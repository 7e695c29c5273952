"""
CPU: host-side logic (no GPU needed) -- the CSR container's dtype rules, COO
construction, sharding/assembly, kernel selection, and that the C-ABI library
loads and exports every symbol include/csrk.h declares.
"""

import os
import pickle
import re

import numpy as np
import pytest

import csr_b200
from csr_b200 import CSR
from csr_b200 import kernels as KS
from csr_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand(rng, nrows, ncols, nnz, dtype='f8', values=True):
    coords = rng.choice(nrows * ncols, nnz, replace=False)
    rows, cols = (coords % nrows).astype(np.int32), (coords // nrows).astype(np.int32)
    vals = rng.normal(size=nnz).astype(dtype) if values else None
    return CSR.from_coo(rows, cols, vals, (nrows, ncols))


# ------------------------------------------------------------------ C ABI
def header_symbols():
    text = open(os.path.join(ROOT, "include", "csrk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csrk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(N.SIGNATURES)


def test_library_exports_every_symbol():
    L = N.lib()            # loads without a GPU; raises if a symbol is missing
    for name in header_symbols():
        assert hasattr(L, name), name
    assert L.csrk_version() >= 100


def test_no_gpu_fails_loudly():
    "Without a device the product path raises; it never falls back to the CPU."
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    K = KS.get_kernel('cuda')
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.to_handle(CSR.empty(1, 1))


def test_product_code_does_not_touch_the_oracle():
    "Only tests/, smoke() and bench.py may reference oracle/ (task rule 3)."
    pkg = os.path.join(ROOT, "csr_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.lower(), os.path.join(dp, fn)


# -------------------------------------------------------- kernel selection
def test_default_kernel_is_cuda():
    "tests/test_active_kernel.py:6-26 of the reference, for the one kernel shipped here"
    assert KS.get_kernel().__name__ == 'csr_b200.kernels.cuda'
    assert KS.get_kernel('cuda').__name__ == 'csr_b200.kernels.cuda'
    from csr_b200 import kernel as static
    assert static.name == 'csr_b200.kernels.cuda'
    for attr in ('to_handle', 'from_handle', 'release_handle', 'order_columns', 'mult_ab', 'mult_abt', 'mult_vec'):
        assert callable(getattr(static, attr))
    assert isinstance(KS.get_kernel().max_nnz, (int, np.integer))


def test_other_kernels_are_not_shipped():
    for name in ('numba', 'mkl', 'scipy'):
        with pytest.raises(ImportError):
            KS.get_kernel(name)


def test_use_kernel_restores():
    "tests/test_active_kernel.py:39-45, plus the nesting the reference gets wrong"
    k0 = KS.get_kernel()
    with KS.use_kernel('cuda'):
        assert KS.get_kernel().__name__ == 'csr_b200.kernels.cuda'
        with KS.use_kernel('cuda'):
            pass
        assert KS._active._active is not None      # still inside the outer block
    assert KS._active._active is None
    assert KS.get_kernel() is k0


def test_selection_is_thread_local():
    import threading
    seen = {}
    KS.set_kernel('cuda')
    try:
        def other():
            seen['active'] = KS._active._active
        t = threading.Thread(target=other)
        t.start()
        t.join()
        assert seen['active'] is None
        assert KS._active._active is not None
    finally:
        KS.set_kernel(None)


def test_releasing_always_releases():
    class K:
        released = []

        @staticmethod
        def release_handle(h):
            K.released.append(h)
    with pytest.raises(RuntimeError):
        with KS.releasing('h1', K) as h:
            assert h == 'h1'
            raise RuntimeError('boom')
    assert K.released == ['h1']


# ------------------------------------------------------------ CSR container
def test_dtype_rules():
    "csr/csr.py:79-100"
    m = CSR(2, 3, 2, np.array([0, 1, 2], np.int64), np.array([0, 2], np.int64), np.array([1, 2], np.float32))
    assert m.rowptrs.dtype == np.int32 and m.colinds.dtype == np.int32 and m.values.dtype == np.float32
    m = CSR(2, 3, 2, np.array([0, 1, 2], np.int64), np.array([0, 2]), None, _cast=False)
    assert m.rowptrs.dtype == np.int64 and m.values is None
    with pytest.raises(AssertionError):
        CSR(-1, 3, 0, np.zeros(1), np.zeros(0), None)


def test_from_coo_keeps_order_within_rows():
    m = CSR.from_coo(np.array([1, 0, 1, 0]), np.array([3, 2, 0, 1]), np.array([1., 2., 3., 4.]), (2, 4))
    assert list(m.rowptrs) == [0, 2, 4]
    assert list(m.colinds) == [2, 1, 3, 0]
    assert list(m.values) == [2., 4., 1., 3.]


def test_known_answer_from_reference_tests():
    "tests/test_transpose.py:11-27 builds this matrix"
    m = CSR.from_coo(np.array([0, 0, 1, 3]), np.array([1, 2, 0, 1]), np.array([0, 1, 2, 3], np.float64), (4, 3))
    assert list(m.rowptrs) == [0, 2, 3, 3, 4]
    assert list(m.row_cs(0)) == [1, 2] and list(m.row_vs(0)) == [0.0, 1.0]
    assert np.array_equal(m.to_scipy().toarray(), [[0, 0, 1], [2, 0, 0], [0, 0, 0], [0, 3, 0]])


def test_shard_and_assemble_roundtrip():
    "tests/test_transform.py:172-197"
    rng = np.random.default_rng(7)
    for _ in range(10):
        m = rand(rng, 60, 40, 500)
        lim = int(rng.integers(int(np.diff(m.rowptrs).max()), 300))
        shards = m._shard_rows(lim)
        assert all(s.nnz <= lim for s in shards)
        assert sum(s.nrows for s in shards) == m.nrows
        m2 = CSR._assemble_shards(shards)
        assert m2.nnz == m.nnz
        assert np.array_equal(m2.rowptrs, m.rowptrs)
        assert np.array_equal(m2.colinds, m.colinds)
        assert np.array_equal(m2.values, m.values)
    with pytest.raises(ValueError):
        m._shard_rows(1)


def test_subset_rows_shares_storage():
    rng = np.random.default_rng(3)
    m = rand(rng, 30, 20, 200)
    s = m.subset_rows(5, 17)
    assert s.nrows == 12 and s.rowptrs[0] == 0
    assert np.shares_memory(s.colinds, m.colinds)
    assert s.nnz == m.rowptrs[17] - m.rowptrs[5]


def test_filter_zeros_host():
    m = CSR(3, 3, 5, np.array([0, 2, 2, 5]), np.array([0, 1, 0, 1, 2]), np.array([1., 0., 0., 2., 0.]))
    m._filter_zeros()
    assert m.nnz == 2 and list(m.rowptrs) == [0, 1, 1, 2]
    assert list(m.colinds) == [0, 1] and list(m.values) == [1., 2.]


def test_pickle_and_copy():
    rng = np.random.default_rng(5)
    m = rand(rng, 10, 12, 40, dtype='f4')
    m2 = pickle.loads(pickle.dumps(m))
    assert (m2.nrows, m2.ncols, m2.nnz) == (m.nrows, m.ncols, m.nnz)
    assert np.array_equal(m2.values, m.values) and m2.values.dtype == np.float32
    c = m.copy()
    c.values[0] = 99
    assert m.values[0] != 99
    assert m.copy(False).values is None


def test_values_setter():
    m = CSR.empty(2, 2, [1, 1])
    m.values = np.array([3., 4., 5.])
    assert list(m.values) == [3., 4.]
    with pytest.raises(ValueError):
        m.values = np.array([1.])
    m.values = None
    assert m.values is None


def test_package_exports():
    assert csr_b200.CSR is CSR
    for n in ('get_kernel', 'set_kernel', 'use_kernel', 'releasing'):
        assert hasattr(csr_b200, n)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: stdout carries ONE JSON line (library banners, logs -> stderr).  The reference arm
    runs on the CPU (the oracle port of the numba kernel), so this is checkable without a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--scale", "0.005"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "spmv_hbm_gbs" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("port", "reference")


def test_every_option_is_documented_in_the_header():
    "csrk_set_option's names (csrc/context.cu) all appear in the tunables comment of include/csrk.h"
    import re
    src = open(os.path.join(ROOT, "csr_b200", "csrc", "context.cu")).read()
    hdr = open(os.path.join(ROOT, "include", "csrk.h")).read()
    names = re.findall(r'!strcmp\(name, "([a-z_0-9]+)"\)', src)
    assert len(names) > 20
    missing = [n for n in names if f'"{n}"' not in hdr]
    assert not missing, missing
