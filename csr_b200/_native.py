"""
ctypes binding of ``libcsr_cuda.so`` (C ABI declared in ``include/csrk.h``).

This is the only place the shared library is loaded.  There is no CPU fallback:
if the library is missing the import fails, and if no B200 is visible the first
call that needs the device raises ``RuntimeError``.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC_DIR, "libcsr_cuda.so")

OK, EARG, ENOMEM, ECUDA, ENODEV, EOVERFLOW = range(6)

# every symbol include/csrk.h declares, with its signature
_vp, _i32, _i64, _int = C.c_void_p, C.c_int32, C.c_int64, C.c_int
_P = C.POINTER
SIGNATURES = {
    "csrk_version": (_int, []),
    "csrk_last_error": (C.c_char_p, []),
    "csrk_init": (_int, [_int]),
    "csrk_shutdown": (_int, []),
    "csrk_device_info": (_int, [_P(_int), _P(_i64), _P(_i64), _P(_int), _P(_int)]),
    "csrk_launch_count": (_i64, []),
    "csrk_synchronize": (_int, []),
    "csrk_set_option": (_int, [C.c_char_p, _i64]),
    "csrk_get_stream": (_int, [_P(_vp)]),
    "csrk_create": (_int, [_i32, _i32, _i64, _vp, _int, _vp, _vp, _int, _P(_vp)]),
    "csrk_create_dev": (_int, [_i32, _i32, _i64, _vp, _int, _vp, _vp, _int, _vp, _P(_vp)]),
    "csrk_free": (_int, [_vp]),
    "csrk_dims": (_int, [_vp, _P(_i32), _P(_i32), _P(_i64), _P(_int), _P(_int)]),
    "csrk_export": (_int, [_vp, _vp, _vp, _vp]),
    "csrk_device_ptrs": (_int, [_vp, _P(_vp), _P(_vp), _P(_vp)]),
    "csrk_subset_rows": (_int, [_vp, _i32, _i32, _P(_vp)]),
    "csrk_spmv": (_int, [_vp, _vp, _int, _vp]),
    "csrk_spmv_plan_info": (_int, [_vp, _int, _P(_i64)]),
    "csrk_spmv_dev": (_int, [_vp, _vp, _int, _vp, _vp]),
    "csrk_spmv_dev_multi": (_int, [_vp, _vp, _int, _P(_vp), _int, _vp]),
    "csrk_spmv_dev_mc": (_int, [_vp, _vp, _int, _vp, _vp, _vp]),
    "csrk_mc_broadcast": (_int, [_vp, _vp, _i64, _vp]),
    "csrk_spgemm": (_int, [_vp, _vp, _P(_vp)]),
    "csrk_spgemm_abt": (_int, [_vp, _vp, _P(_vp)]),
    "csrk_spgemm_stats": (_int, [_vp, _P(_i64), _P(_i64)]),
    "csrk_spgemm_path": (_int, [_vp, _P(_int)]),
    "csrk_spgemm_side_list": (_int, [_vp, _P(_i64)]),
    "csrk_normalize_rows": (_int, [_vp, _int, _vp, _vp]),
    "csrk_from_coo": (_int, [_i32, _i32, _i64, _vp, _vp, _vp, _int, _P(_vp)]),
    "csrk_transpose": (_int, [_vp, _int, _P(_vp)]),
    "csrk_order_columns": (_int, [_vp]),
    "csrk_filter_zeros": (_int, [_vp]),
}

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (``csrc/Makefile``)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "-s", "clean"], check=True)
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libcsr_cuda.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


def lib():
    """The loaded library (loads on first use; never falls back to anything else)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` "
                "(or __graft_entry__.build()).  csr_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    msg = lib().csrk_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    """Map a csrk status to the Python exception the reference's callers expect
    (ValueError for bad shapes/capacity: csr/kernels/mkl/handle.py:62-63)."""
    if rc == OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == EARG:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == EOVERFLOW:
        raise OverflowError(msg)
    raise RuntimeError(msg)


def device_info() -> dict:
    sm, tot, free, maj, mnr = _int(), _i64(), _i64(), _int(), _int()
    check(lib().csrk_device_info(C.byref(sm), C.byref(tot), C.byref(free), C.byref(maj), C.byref(mnr)), "device_info")
    return {"sm_count": sm.value, "mem_total": tot.value, "mem_free": free.value, "cc": (maj.value, mnr.value)}


def launch_count() -> int:
    return int(lib().csrk_launch_count())
