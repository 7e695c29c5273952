import logging
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hypothesis import HealthCheck, Phase, settings  # noqa: E402

logging.getLogger('numba').setLevel(logging.INFO)

# hypothesis profiles, as the reference's conftest.py:41-47
# no shrink phase: a failing draw is reported as drawn (shrinking through a GPU call took minutes)
settings.register_profile('default', deadline=None, max_examples=60,
                          suppress_health_check=list(HealthCheck),
                          phases=[Phase.explicit, Phase.reuse, Phase.generate, Phase.target])
settings.register_profile('large', settings.get_profile('default'), max_examples=2000)
settings.register_profile('fast', settings.get_profile('default'), max_examples=20)
settings.load_profile(os.environ.get('HYPOTHESIS_PROFILE', 'default'))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    "Outputs of the unmodified reference (tests/golden/make_golden.py)."
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    return z


@pytest.fixture(scope="module")
def kernel():
    """
    The kernel under test, selected the way the reference's ``kernel`` fixture does
    (conftest.py:20-37): ``use_kernel(name)`` + a warm-up handle on a 1x1 matrix.
    Only ``cuda`` exists here; it needs a GPU.
    """
    from csr_b200 import CSR
    from csr_b200.kernels import use_kernel, get_kernel
    with use_kernel('cuda'):
        k = get_kernel()
        m = CSR.empty(1, 1)
        h = k.to_handle(m)
        k.release_handle(h)
        del h, m
        yield k
