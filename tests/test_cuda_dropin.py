"""
Drop-in proof: the UNMODIFIED reference package (staged by ``oracle/stage_ref.py`` into the git-ignored
``oracle/_ref/``) drives the ``cuda`` kernel through its own selection API, its own ``CSR.multiply`` /
``CSR.mult_vec`` (csr/csr.py:524-590), its own ``kernel`` fixture (conftest.py:20-37, with "cuda" appended
to ``KERNELS``: the one-line test integration) and its own hot-path tests.

The only file added to the reference tree is ``csr/kernels/cuda/__init__.py``
(``csr_b200/integration/csr_kernels_cuda.py``).

Every check runs in a subprocess: the reference package is called ``csr`` and reads ``CSR_KERNEL`` at its
first ``get_kernel()``, so it gets an interpreter of its own.
"""

import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(STAGE, "csr")),
                                reason="reference not staged (run oracle/stage_ref.py in the build container)")


def run_py(code=None, args=None, env_extra=None, timeout=1500, cwd=None):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([STAGE, ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    env.pop("CSR_KERNEL", None)
    env.update(env_extra or {})
    cmd = [sys.executable] + (["-c", code] if code is not None else list(args))
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=cwd or STAGE)


def ok(r):
    assert r.returncode == 0, f"rc {r.returncode}\n--- stdout\n{r.stdout[-4000:]}\n--- stderr\n{r.stderr[-4000:]}"
    return r


# ------------------------------------------------------------------ no GPU needed
def test_reference_selection_api_resolves_cuda():
    "get_kernel('cuda') / use_kernel('cuda') of the reference's own selection module (csr/kernels/__init__.py:66-97)."
    r = ok(run_py("""
import json, csr
from csr.kernels import get_kernel, use_kernel, set_kernel
K = get_kernel('cuda')
assert K.__name__ == 'csr.kernels.cuda', K.__name__
need = ['max_nnz', 'to_handle', 'from_handle', 'release_handle', 'order_columns', 'mult_ab', 'mult_abt', 'mult_vec']
missing = [n for n in need if not hasattr(K, n)]
assert not missing, missing
default = get_kernel()
assert default.__name__ == 'csr.kernels.numba', default.__name__   # cuda is never the silent default
with use_kernel('cuda'):
    assert get_kernel() is K
assert get_kernel() is default
set_kernel('cuda'); assert get_kernel() is K; set_kernel(None); assert get_kernel() is default
print(json.dumps({'file': K.__file__}))
"""))
    assert json.loads(r.stdout.strip().splitlines()[-1])["file"].startswith(STAGE)


def test_reference_env_selects_cuda_and_static_kernel_binds_it():
    "CSR_KERNEL=cuda (csr/kernels/__init__.py:107-109) and the static csr.kernel module (csr/kernel.py:5-16)."
    ok(run_py("""
import csr, csr.kernel
from csr.kernels import get_kernel
assert get_kernel().__name__ == 'csr.kernels.cuda'
assert csr.kernel.name == 'csr.kernels.cuda'
import csr_b200.kernels.cuda as ours
assert csr.kernel.mult_vec is ours.mult_vec and csr.kernel.mult_abt is ours.mult_abt
""", env_extra={"CSR_KERNEL": "cuda"}))


# ------------------------------------------------------------------ on the GPU box
REF_TESTS = ["test_handles.py", "test_mult_vec.py", "test_multiply.py", "test_transform.py"]


@pytest.mark.gpu
def test_reference_hot_path_tests_pass_under_cuda():
    """The reference's own tests that take the ``kernel`` fixture (tests/test_handles.py:10-21,
    test_mult_vec.py:12-39, test_multiply.py:14-79, test_transform.py:77-87), selected with -k cuda."""
    r = ok(run_py(args=["-m", "pytest", "-q", "-p", "no:cacheprovider", "-k", "cuda", "-rA"] + REF_TESTS,
                  cwd=os.path.join(STAGE, "reftests")))
    passed = set(re.findall(r"^PASSED (\S+)", r.stdout, flags=re.M))
    want = {"test_handles.py::test_make_handle[cuda]", "test_mult_vec.py::test_mult_vec[cuda]",
            "test_mult_vec.py::test_mult_vec_novalue[cuda]", "test_multiply.py::test_multiply[cuda]",
            "test_multiply.py::test_multiply_transpose[cuda]", "test_transform.py::test_kernel_sort_rows[cuda]"}
    assert want <= passed, f"missing: {sorted(want - passed)}\n{r.stdout[-3000:]}"
    assert " failed" not in r.stdout.splitlines()[-1]


@pytest.mark.gpu
def test_reference_sharding_path_under_cuda():
    """The ``max_nnz`` trick of tests/test_mkl.py:29-38,76-91: lower the kernel's capacity so the reference's
    ``CSR.multiply`` / ``CSR.mult_vec`` go through ``_shard_rows`` + ``_assemble_shards`` (csr.py:599-650)."""
    ok(run_py("""
import sys
sys.path.insert(0, '.')
from csr.kernels import use_kernel, get_kernel
import test_multiply as tmm, test_mult_vec as tmv
cuda = get_kernel('cuda')
save = cuda.max_nnz
cuda.max_nnz = 1000
try:
    with use_kernel('cuda'):
        tmv.test_mult_vec(cuda)
        tmm.test_multiply(cuda)
        tmm.test_multiply_transpose(cuda)
finally:
    cuda.max_nnz = save
""", cwd=os.path.join(STAGE, "reftests")))


@pytest.mark.gpu
def test_reference_csr_object_api_matches_numba_kernel():
    """``CSR_KERNEL=cuda``: the reference's ``CSR.multiply`` / ``CSR.mult_vec`` / kernel-level calls on the
    reference's own ``CSR`` objects against the reference's numba kernel in the same process: rowptrs
    bit-exact, colinds bit-exact after the canonical sort, values rtol 1e-10."""
    ok(run_py("""
import numpy as np, scipy.sparse as sps
import csr
from csr import CSR
from csr.kernels import get_kernel, use_kernel
assert get_kernel().__name__ == 'csr.kernels.cuda'
rng = np.random.default_rng(5)
def rand(nr, nc, dens, dt='f8'):
    m = sps.random(nr, nc, dens, format='csr', random_state=rng, data_rvs=lambda n: rng.uniform(0.5, 5.0, n)).astype(dt)
    return CSR.from_scipy(m)
A, B, Bt = rand(300, 200, 0.05), rand(200, 150, 0.08), rand(170, 200, 0.06, 'f4')
x = rng.standard_normal(200)
def canon(m):
    rows = np.repeat(np.arange(m.nrows), np.diff(m.rowptrs))
    o = np.lexsort((m.colinds, rows))
    return m.rowptrs, m.colinds[o], m.values[o]
for (lhs, rhs, tr) in ((A, B, False), (A, Bt, True)):
    got = lhs.multiply(rhs, transpose=tr)
    assert type(got) is CSR
    with use_kernel('numba'):
        ref = lhs.multiply(rhs, transpose=tr)
    rp, ci, vs = canon(ref)
    assert got.rowptrs.dtype == rp.dtype and np.array_equal(got.rowptrs, rp)
    assert np.array_equal(got.colinds, ci)
    assert np.allclose(got.values, vs, rtol=1e-5 if tr else 1e-10, atol=0)
y = A.mult_vec(x)
with use_kernel('numba'):
    y0 = A.mult_vec(x)
assert y.dtype == np.float64 and np.allclose(y, y0, rtol=1e-10, atol=1e-12)
# kernel-level lifecycle on the reference's objects (tests/test_handles.py:10-21)
K = get_kernel()
h = K.to_handle(A); c = K.from_handle(h); K.release_handle(h); K.release_handle(h)
assert type(c) is CSR and (c.nrows, c.ncols, c.nnz) == (A.nrows, A.ncols, A.nnz)
assert np.array_equal(c.rowptrs, A.rowptrs) and np.array_equal(c.colinds, A.colinds) and np.array_equal(c.values, A.values)
""", env_extra={"CSR_KERNEL": "cuda"}))


@pytest.mark.gpu
def test_reference_nopython_callers_reach_cuda():
    """``CSR_KERNEL=cuda`` serves the STATIC kernel path too: ``csr/kernel.py:9-16`` binds the cuda functions and the
    ``@overload_method``s of csr/_wiring.py:116-151 call them from ``@njit`` code on ``CSRType`` structrefs
    (to_handle -> mult_ab / mult_abt / mult_vec -> from_handle -> release_handle, all nopython).  Results against the
    numba kernel in object mode: rowptrs bit-exact, colinds after the canonical sort, values rtol 1e-10."""
    ok(run_py("""
import numpy as np, scipy.sparse as sps
from numba import njit
import csr, csr.kernel
from csr import CSR
from csr.kernels import use_kernel
assert csr.kernel.name == 'csr.kernels.cuda'
rng = np.random.default_rng(9)
def rand(nr, nc, dens):
    m = sps.random(nr, nc, dens, format='csr', random_state=rng, data_rvs=lambda n: rng.uniform(0.5, 5.0, n))
    return CSR.from_scipy(m)
A, B, Bt = rand(250, 180, 0.06), rand(180, 140, 0.07), rand(160, 180, 0.05)
x = rng.standard_normal(180)

@njit
def products(a, b, bt, x):
    return a.multiply(b, False), a.multiply(bt, True), a.mult_vec(x)

ab, abt, y = products(A, B, Bt, x)
def canon(m):
    rows = np.repeat(np.arange(m.nrows), np.diff(m.rowptrs))
    o = np.lexsort((m.colinds, rows))
    return m.rowptrs, m.colinds[o], m.values[o]
with use_kernel('numba'):
    r_ab, r_abt, r_y = A.multiply(B), A.multiply(Bt, transpose=True), A.mult_vec(x)
for got, ref in ((ab, r_ab), (abt, r_abt)):
    assert type(got) is CSR
    rp, ci, vs = canon(ref)
    assert got.rowptrs.dtype == rp.dtype and np.array_equal(got.rowptrs, rp) and np.array_equal(got.colinds, ci)
    assert np.allclose(got.values, vs, rtol=1e-10, atol=0)
assert np.allclose(y, r_y, rtol=1e-10, atol=1e-12)
from csr_b200 import _native
assert _native.launch_count() > 0          # the products above ran on the GPU
""", env_extra={"CSR_KERNEL": "cuda"}))
