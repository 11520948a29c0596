"""
ngx_http_imgproc_b200 — B200 (sm_100a) implementation of the decoded-pixel hot path of
tommiv/ngx_http_imgproc: crop, resize, filter chain, watermark compositing, alpha flatten
(RunJob steps 3-7, bridge.c:574-656), behind the C ABI in include/imp_gpu.h.

Layout:
  csrc/      CUDA kernels, host planner (the reference's argument grammar), C-ABI runtime
  build.py   nvcc build of libimp_gpu.so (in-tree)
  api.py     ctypes binding used by tests/ and bench.py

There is no CPU implementation in this package. Importing it does not need a GPU; computing does.
"""
from . import build as build_mod
from .api import (Batch, Config, ImpError, Library, Plan, library, run_host_batch,  # noqa: F401
                  IMP_OK, IMP_ERROR_GPU, IMP_ERROR_INVALID_ARGS, IMP_ERROR_NO_SUCH_FILTER,
                  IMP_ERROR_TOO_BIG_TARGET, IMP_ERROR_TOO_MUCH_FILTERS, INTERP_LINEAR, INTERP_REFERENCE)


def build(force: bool = False, verbose: bool = False) -> str:
    return build_mod.build(force=force, verbose=verbose)
