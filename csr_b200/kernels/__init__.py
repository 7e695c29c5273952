"""
Kernel selection -- the host-side mirror of ``csr/kernels/__init__.py``
(reference lines 1-123): a name -> module registry filled by ``import_module``,
a THREAD-LOCAL active kernel, the ``CSR_KERNEL`` environment override, and the
``releasing`` context manager that guarantees ``release_handle``.

This package ships exactly one kernel, ``cuda``; it is the default.  There is no
multi-backend dispatch and no CPU fallback: asking for any other name raises
``ImportError`` exactly as the reference does for a kernel that is not installed.

One deliberate fix: the reference's ``use_kernel`` restores ``active_name``,
which it never updates (csr/kernels/__init__.py:18,73), so leaving a nested block
always falls back to the default.  Here the previously active kernel is restored;
the un-nested behaviour the reference tests (tests/test_active_kernel.py:39-45)
is unchanged.
"""

import os
import threading
import warnings
from contextlib import contextmanager
from importlib import import_module

kernels = {}
__all__ = [
    'releasing',
    'set_kernel',
    'use_kernel',
    'get_kernel',
]

DEFAULT_KERNEL = 'cuda'


class ActiveKernel(threading.local):
    "Thread-local slot for the explicitly selected kernel (reference :16-29)."

    def __init__(self):
        self._active = None

    @property
    def active(self):
        kern = self._active
        if kern is None:
            return _default_kernel()
        return kern

    def set_active(self, kern):
        self._active = kern


_cached_default = None
_active = ActiveKernel()


@contextmanager
def releasing(h, k):
    "Yield ``h`` and release it through kernel ``k`` on exit (reference :36-41)."
    try:
        yield h
    finally:
        k.release_handle(h)


def set_kernel(name):
    """
    Set the (thread-local) active kernel; ``None`` returns to the default
    (reference :44-63).  Does not change the statically bound ``csr_b200.kernel``.
    """
    if name is None:
        _active.set_active(None)
    else:
        _active.set_active(get_kernel(name))


@contextmanager
def use_kernel(name):
    "Run a block with the named kernel active, then restore the previous one (reference :66-78)."
    old = _active._active
    try:
        set_kernel(name)
        yield
    finally:
        _active.set_active(old)


def get_kernel(name=None):
    "The named kernel module, or the active one when ``name`` is None (reference :81-97)."
    if name is None:
        return _active.active

    kern = kernels.get(name, None)
    if not kern:
        kern = import_module(f'{__name__}.{name}')
        kernels[name] = kern
    return kern


def _initialize(name=None):
    "Pick the process default: explicit name, else $CSR_KERNEL, else cuda (reference :100-116)."
    global _cached_default
    if _cached_default:
        warnings.warn('default kernel already initialized')

    if not name:
        name = os.environ.get('CSR_KERNEL', DEFAULT_KERNEL)
    _cached_default = import_module(f'{__name__}.{name}')


def _default_kernel():
    if not _cached_default:
        _initialize()
    return _cached_default
