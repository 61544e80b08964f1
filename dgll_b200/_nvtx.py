"""NVTX ranges with the names the reference gives its profiler ranges (``torch.autograd.profiler.record_function`` at
dgll/FeatureCache/storage.py:164-206: cache-idxload / cache-index / cache-allocate / cache-gpu / cache-cpu / cache-asign;
dgll/FeatureCache/gcn.py:83-93: gpu-load / gpu-compute), so an Nsight timeline of this path lines up with one of the
reference.  Off by default; on with the library option ``nvtx`` (``kernels.set_option("nvtx", 1)`` or DGLLB_NVTX=1)."""
import contextlib

import torch

from . import _lib


def enabled():
    return _lib.get_option("nvtx") == 1


@contextlib.contextmanager
def range(name):
    if not enabled():
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
