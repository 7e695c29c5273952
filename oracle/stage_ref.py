#!/usr/bin/env python
"""
Stage the UNMODIFIED reference package next to the oracle, in git-ignored ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  The reference (lenskit/csr) is pure Python + numba; it cannot be pip-installed
offline (its build backend, flit_core, is not in the image), so "building" it is a copy of the package
directory.  ``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never enter the history) but
not in ``.gpurunignore``, so the staged tree travels to the GPU box like the built ``.so`` files.  There it
provides (a) the reference's own numba kernel as the CPU arm of ``bench.py`` (``cpu_baseline.kind ==
"reference"``) and (b) the reference's own ``CSR`` class, kernel-selection module, ``kernel`` fixture and
hot-path tests, which ``tests/test_cuda_dropin.py`` runs against the ``cuda`` kernel.

What is written:
    oracle/_ref/csr/                       copy of /root/reference/csr
    oracle/_ref/csr/kernels/cuda/__init__.py   OUR file: csr_b200/integration/csr_kernels_cuda.py
    oracle/_ref/reftests/conftest.py       the reference's conftest.py with "cuda" appended to KERNELS
                                           (the one-line test integration of SURVEY.md section 4)
    oracle/_ref/reftests/test_*.py         the reference's hot-path test files, unmodified
    oracle/_ref/STAGED.json                what was copied, from where

Run by ``__graft_entry__.build()`` when /root/reference is present (the build container); the GPU box only
uses the staged files.
"""

import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEST = os.path.join(HERE, "_ref")
REF = os.environ.get("CSR_REFERENCE", "/root/reference")

# the reference's tests that reach the kernel hot path
HOT_TESTS = ["test_handles.py", "test_mult_vec.py", "test_multiply.py", "test_transform.py",
             "test_active_kernel.py", "test_kernel_numba.py", "test_numba.py"]


def stage(force=False):
    src_pkg = os.path.join(REF, "csr")
    if not os.path.isdir(src_pkg):
        return False
    pkg = os.path.join(DEST, "csr")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(src_pkg, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    # the maintainer's one new file
    os.makedirs(os.path.join(pkg, "kernels", "cuda"), exist_ok=True)
    shutil.copyfile(os.path.join(ROOT, "csr_b200", "integration", "csr_kernels_cuda.py"),
                    os.path.join(pkg, "kernels", "cuda", "__init__.py"))
    # the reference's own tests + fixture, with the cuda kernel added to the fixture's list
    tdir = os.path.join(DEST, "reftests")
    os.makedirs(tdir)
    conf = open(os.path.join(REF, "conftest.py")).read()
    conf2, n = re.subn(r'^KERNELS = \[(.*)\]$', r'KERNELS = [\1, "cuda"]', conf, count=1, flags=re.M)
    assert n == 1, "conftest.py: KERNELS list not found"
    open(os.path.join(tdir, "conftest.py"), "w").write(conf2)
    copied = []
    for t in HOT_TESTS:
        p = os.path.join(REF, "tests", t)
        if os.path.exists(p):
            shutil.copyfile(p, os.path.join(tdir, t))
            copied.append(t)
    # pytest.ini of the reference asks for pytest-benchmark options (absent here): keep only its filters
    open(os.path.join(tdir, "pytest.ini"), "w").write(
        "[pytest]\nfilterwarnings =\n    ignore:.*matrix subclass.*:PendingDeprecationWarning\n"
        "    ignore:.*is a deprecated alias.*:DeprecationWarning\n    ignore:.*use CSR directly.*:DeprecationWarning\n")
    json.dump({"reference": REF, "package": "csr", "tests": copied,
               "conftest_change": 'KERNELS += ["cuda"]',
               "added": "csr/kernels/cuda/__init__.py <- csr_b200/integration/csr_kernels_cuda.py"},
              open(os.path.join(DEST, "STAGED.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else f"{REF} not present: nothing staged", file=sys.stderr)
