"""Reference-side integration files: what a lenskit/csr maintainer adds to use this package
(INTEGRATION.md section 1).  ``csr_kernels_cuda.py`` is installed as ``csr/kernels/cuda/__init__.py``."""
