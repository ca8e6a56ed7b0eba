"""Process-wide handles (reference smart_compress/util/globals.py:5-7).

``profiler`` is anything with a ``profile(name)`` context manager (the reference stores
Lightning's profiler here, util/train.py:217).  ``None`` means "no profiling"; ``NvtxProfiler``
below turns each plugin call into an NVTX range that ncu / Nsight can filter on.
"""
import contextlib


class NvtxProfiler:
    @contextlib.contextmanager
    def profile(self, name):
        import torch

        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()


class Globals:
    compression = None
    profiler = None
