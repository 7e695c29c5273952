"""
The UNMODIFIED reference's numba kernel (staged in ``oracle/_ref`` by ``oracle/stage_ref.py``) as a timed
CPU baseline.  TEST/BENCH INFRASTRUCTURE ONLY: imported by ``bench.py``'s ``cpu_baseline`` and
``--impl reference`` legs and by tests, never by ``csr_b200``.

``available()`` says whether the staged package and numba import here; ``load()`` returns the reference's
``csr`` package with the numba kernel active.  The kernels are serial as shipped (no ``prange`` anywhere:
csr/kernels/numba/__init__.py:55, multiply.py:13,41,60,103) and ``nogil``, so the all-cores figures run a
``ThreadPoolExecutor`` over the reference's own ``CSR._shard_rows`` (csr/csr.py:599-621) without touching
reference code.
"""

import math
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGE = os.path.join(_HERE, "_ref")
_mod = None
_err = None


def load():
    "The staged reference package ``csr`` (numba kernel as the process default)."
    global _mod, _err
    if _mod is None and _err is None:
        if not os.path.isdir(os.path.join(STAGE, "csr")):
            _err = "oracle/_ref/csr is not staged"
            return None
        saved_env, saved_path, saved_flag = os.environ.get("CSR_KERNEL"), sys.path[:], sys.dont_write_bytecode
        os.environ["CSR_KERNEL"] = "numba"
        sys.dont_write_bytecode = True
        sys.path.insert(0, STAGE)
        try:
            import csr as ref
            from csr.kernels import get_kernel
            assert get_kernel().__name__ == "csr.kernels.numba"
            _mod = ref
        except Exception as e:  # numba missing, ...
            _err = repr(e)
        finally:
            sys.path[:] = saved_path
            sys.dont_write_bytecode = saved_flag
            if saved_env is None:
                os.environ.pop("CSR_KERNEL", None)
            else:
                os.environ["CSR_KERNEL"] = saved_env
    return _mod


def available():
    return load() is not None


def why_not():
    return _err


def kernel():
    from csr.kernels import get_kernel   # noqa: resolved through sys.modules after load()
    return get_kernel("numba")


def as_ref(m):
    "Any six-field CSR record as the reference's own ``csr.CSR`` (no copy for matching dtypes)."
    ref = load()
    return ref.CSR(m.nrows, m.ncols, m.nnz, np.asarray(m.rowptrs), np.asarray(m.colinds),
                   None if m.values is None else np.asarray(m.values))


def shards(A, threads):
    "The reference's own row sharding, sized for `threads` workers."
    if threads <= 1:
        return [A]
    longest = int(np.diff(A.rowptrs).max(initial=1))
    return A._shard_rows(max(int(math.ceil(A.nnz / threads)), longest, 1))


def mult_vec(A, x, parts=None, pool=None):
    "numba mult_vec on one core (``parts`` None) or over row shards on a thread pool."
    K = kernel()
    if parts is None:
        return K.mult_vec(A, x)
    return np.concatenate(list(pool.map(lambda s: K.mult_vec(s, x), parts)))


def mult_ab_parts(parts, B, pool, transpose=False):
    "numba mult_ab / mult_abt of each row shard against B on a thread pool (blocks are returned unassembled)."
    K = kernel()
    if transpose:
        B = B.transpose()
    return list(pool.map(lambda s: K.mult_ab(s, B), parts))


def pool(threads):
    return ThreadPoolExecutor(max_workers=max(threads, 1))
