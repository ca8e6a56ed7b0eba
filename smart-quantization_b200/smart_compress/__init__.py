"""Host-side mirror of the reference's `smart_compress` plugin surface for the
compress->decompress hot path, backed by hand-written sm_100a CUDA (libsmaq_b200.so).

Module paths, class names, flag names and defaults follow the reference
(nimashoghi/smart-quantization, `smart_compress/`), so code written against the reference's
compression classes and hooks imports this package unchanged.  Only the hot path is here:
trainer, datasets and model zoo are out of scope.
"""
__all__ = ["compress", "util"]
