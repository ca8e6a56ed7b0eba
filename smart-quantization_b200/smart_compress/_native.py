"""ctypes binding of libsmaq_b200.so — the C ABI declared in include/smaq_b200.h.

This is the only place the host side touches native code.  There is no CPU path and no
fallback: if the library is missing or a call fails, an exception is raised.

Torch is used for what the boundary leaves to the caller: device buffers
(``tensor.data_ptr()``), workspaces and the current CUDA stream of the calling thread
(``torch.cuda.current_stream()``), which is how backward-pass calls made on autograd's worker
thread end up on the right stream.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMAQ_B200_LIB: development override (A/B builds of the same ABI, tools/ab_build.sh); unset in normal use
LIB_PATH = os.environ.get("SMAQ_B200_LIB") or os.path.join(_HERE, "_lib", "libsmaq_b200.so")

OK = 0
ABI_VERSION = 4


class NativeLibraryError(RuntimeError):
    pass


class CodecParams(C.Structure):  # smaq_codec_params
    _fields_ = [
        ("threshold", C.c_float),
        ("range_main", C.c_float),
        ("range_outlier", C.c_float),
        ("clamp_lo", C.c_float),
        ("clamp_hi", C.c_float),
        ("bits_main", C.c_int32),
        ("bits_outlier", C.c_int32),
        ("stochastic", C.c_int32),
        ("all_positive", C.c_int32),
        ("saturate", C.c_int32),
        ("count_saturated", C.c_int32),
        ("zero_on_grid", C.c_int32),
        ("seed", C.c_uint64),
        ("offset", C.c_uint64),
        ("offset_base", C.c_void_p),
    ]


class FloatqParams(C.Structure):  # smaq_floatq_params
    _fields_ = [
        ("exp_bits", C.c_int32),
        ("man_bits", C.c_int32),
        ("rounding", C.c_int32),
        ("check_inf", C.c_int32),
        ("max_exp_bias", C.c_int32),
        ("reserved", C.c_int32),
        ("seed", C.c_uint64),
        ("offset", C.c_uint64),
        ("offset_base", C.c_void_p),
    ]


class TensorDesc(C.Structure):  # smaq_tensor_desc
    _fields_ = [
        ("x", C.c_void_p),
        ("y", C.c_void_p),
        ("n", C.c_int64),
        ("all_positive", C.c_int32),
        ("stream", C.c_int32),
    ]


class PackedLayout(C.Structure):  # smaq_packed_layout
    _fields_ = [
        ("n", C.c_int64),
        ("bits_main", C.c_int32),
        ("bits_outlier", C.c_int32),
        ("n_warp_tiles", C.c_int64),
        ("n_cta_tiles", C.c_int64),
        ("header_off", C.c_int64),
        ("header_bytes", C.c_int64),
        ("planes_off", C.c_int64),
        ("planes_bytes", C.c_int64),
        ("extras_off", C.c_int64),
        ("extras_stride_bytes", C.c_int64),
        ("extras_capacity_bytes", C.c_int64),
        ("total_capacity_bytes", C.c_int64),
        ("workspace_bytes", C.c_int64),
    ]


class PackedHeader(C.Structure):  # smaq_packed_header
    _fields_ = [
        ("magic", C.c_uint32),
        ("bits_main", C.c_int32),
        ("bits_outlier", C.c_int32),
        ("stochastic", C.c_int32),
        ("n", C.c_int64),
        ("mean", C.c_float),
        ("std_raw", C.c_float),
        ("threshold", C.c_float),
        ("range_main", C.c_float),
        ("range_outlier", C.c_float),
        ("clamp_lo", C.c_float),
        ("clamp_hi", C.c_float),
        ("pad0", C.c_float),
        ("n_outlier", C.c_uint64),
        ("n_saturated", C.c_uint64),
        ("extras_words", C.c_uint64),
        ("status", C.c_uint64),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_SIGNATURES = {
    # name: (restype, argtypes)
    "smaq_b200_abi_version": (C.c_int, []),
    "smaq_b200_last_error": (C.c_char_p, []),
    "smaq_b200_sm_count": (C.c_int, []),
    "smaq_counter_add": (C.c_int, [_P, C.c_uint64, _P]),
    "smaq_stats_workspace_bytes": (C.c_size_t, [_I64]),
    "smaq_stats_full": (C.c_int, [_P, _I64, C.c_int, _P, _P, C.c_size_t, _P]),
    "smaq_stats_range": (C.c_int, [_P, _I64, _P, _P, C.c_size_t, _P]),
    "smaq_stats_sampled": (C.c_int, [_P, _I64, _P, C.c_int32, C.c_int32, _P, _P]),
    "smaq_stats_sampled_draw": (C.c_int, [_P, _I64, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, _P, _P]),
    "smaq_roundtrip": (C.c_int, [_P, _P, _I64, _P, _P, C.POINTER(CodecParams), _P]),
    "smaq_compress_workspace_bytes": (C.c_size_t, [_I64]),
    "smaq_compress_workspace_init": (C.c_int, [_P, C.c_size_t, _P]),
    "smaq_compress": (C.c_int, [_P, _P, _I64, _P, C.POINTER(CodecParams), _P, C.c_size_t, _P]),
    "smaq_count_outliers": (C.c_int, [_P, _I64, _P, C.POINTER(CodecParams), _P, _P]),
    "smaq_fused_small_max": (_I64, []),
    "smaq_roundtrip_small": (C.c_int, [_P, _P, _I64, _P, C.POINTER(CodecParams), _P, _P]),
    "smaq_multi_workspace_bytes": (C.c_size_t, [C.c_int32, _I64]),
    "smaq_roundtrip_multi": (C.c_int, [_P, C.c_int32, _I64, _I64, C.POINTER(CodecParams), _I64, _P, C.c_size_t, _P, _P]),
    "smaq_roundtrip_bn": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _I64, _I64, C.POINTER(CodecParams), _P]),
    "smaq_packed_layout_for": (C.c_int, [_I64, C.c_int32, C.c_int32, C.POINTER(PackedLayout)]),
    "smaq_encode_workspace_init": (C.c_int, [_P, C.c_size_t, _P]),
    "smaq_encode": (C.c_int, [_P, _I64, _P, _P, C.POINTER(CodecParams), _P, C.c_size_t, _P, C.c_size_t, _P]),
    "smaq_decode": (C.c_int, [_P, C.c_size_t, _I64, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "smaq_encode_split": (C.c_int, [_P, _I64, _P, _P, C.POINTER(CodecParams), _P, C.c_size_t, _P, C.c_size_t, _P, C.c_size_t, _P]),
    "smaq_decode_split": (C.c_int, [_P, C.c_size_t, _P, C.c_size_t, _P, _I64, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "smaq_extras_table_entries": (_I64, [_I64]),
    "smaq_extras_compact": (C.c_int, [_P, C.c_size_t, _P, _I64, C.c_int32, C.c_int32, _P, _P, C.c_size_t, _P, _P]),
    "smaq_decode_sum": (C.c_int, [_P, C.c_int32, C.c_size_t, _I64, C.c_int32, C.c_int32, _I64, _I64, C.c_float, _P, _P]),
    "smaq_float_quantize": (C.c_int, [_P, _P, _I64, _P, C.POINTER(FloatqParams), _P]),
    "smaq_floatq_multi_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "smaq_float_quantize_multi": (C.c_int, [_P, C.c_int32, _I64, C.POINTER(FloatqParams), _P, C.c_size_t, _P]),
    "smaq_s2fp8_multi_workspace_bytes": (C.c_size_t, [C.c_int32, _I64]),
    "smaq_s2fp8_multi": (C.c_int, [_P, C.c_int32, _I64, C.POINTER(FloatqParams), _P, C.c_size_t, _P, _P]),
    "smaq_s2fp8_stats": (C.c_int, [_P, _I64, _P, _P, C.c_size_t, _P]),
    "smaq_s2fp8_apply": (C.c_int, [_P, _P, _I64, _P, _P, C.POINTER(FloatqParams), _P]),
    "smaq_selftest_pow": (C.c_int, [_P, _P, _P, _P, _I64, _P]),
    "smaq_selftest_s2_screen": (C.c_int, [_P, _I64, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """dlopen the library once; raises NativeLibraryError when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python smart-quantization_b200/build.py` "
                "(there is no CPU or eager fallback for this path)"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if lib.smaq_b200_abi_version() != ABI_VERSION:
            raise NativeLibraryError("libsmaq_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != OK:
        msg = load().smaq_b200_last_error()
        raise NativeLibraryError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device=None) -> int:
    """The calling thread's current CUDA stream on ``device`` as an integer handle (2 us through the
    ``torch.cuda.Stream`` object, 0.2 us through the raw getter when this torch build exposes it)."""
    if _raw_stream is not None:
        if device is None:
            index = torch.cuda.current_device()
        elif isinstance(device, int):
            index = device
        else:
            if not isinstance(device, torch.device):
                device = torch.device(device)  # "cuda:0"
            index = device.index if device.index is not None else torch.cuda.current_device()
        return _raw_stream(index)
    return torch.cuda.current_stream(device).cuda_stream


_get_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


class _NoSwitch:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def on_device_of(t: torch.Tensor):
    """Context in which the current CUDA device is ``t``'s.  The C entry points launch on the CURRENT device and
    size their grids from it; a tensor on cuda:1 while cuda:0 is current (``model.to("cuda:1")`` without
    ``set_device``) would otherwise be launched on the wrong GPU.  The reference's eager operators work on any
    device.  Free when the devices already agree (one integer compare); a device switch otherwise."""
    index = t.device.index
    if index is None or index == _get_device():
        return _NO_SWITCH
    return torch.cuda.device(index)


def wrong_device(t: torch.Tensor) -> bool:
    index = t.device.index
    return index is not None and index != _get_device()


# ---- random-stream numbering that survives CUDA-graph replay ---------------------------------------------------
# Every codec call draws its own Philox stream.  Eagerly the stream number is a host counter passed by value; in a
# captured training step that value is baked into the graph, so every replay would round with the same numbers.
# Inside ``counted_step(device)`` the calls are numbered 0, 1, 2 ... from the start of the step and the kernels add
# a DEVICE counter to that number (smaq_codec_params.offset_base); when the step ends the counter is advanced by the
# number of streams used (smaq_counter_add, one thread, enqueued on the current stream — captured with the step).
# Eager execution and graph replay therefore draw the same numbers, step for step.
class StepCounter:
    def __init__(self, device):
        self.base = torch.zeros(1, dtype=torch.int64, device=device)   # read by the kernels as uint64
        self.calls = 0

    def next(self, count: int = 1) -> int:
        first = self.calls
        self.calls += count
        return first

    @property
    def base_ptr(self) -> int:
        return self.base.data_ptr()


_step_counters = {}   # device index -> StepCounter (created once per device: graphs hold its address)
_active_counters = {}  # device index -> StepCounter, while inside counted_step


def active_counter():
    """The StepCounter of the current device if a counted step is open, else None (one dict lookup)."""
    if not _active_counters:
        return None
    return _active_counters.get(_get_device())


class counted_step:
    """``with counted_step(device): loss = opt.step(closure)`` — see above.  Re-entrant per device is an error."""

    def __init__(self, device=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)

    def __enter__(self):
        if self.index in _active_counters:
            raise RuntimeError("counted_step is already open on this device")
        sc = _step_counters.get(self.index)
        if sc is None:
            sc = _step_counters[self.index] = StepCounter(self.device)
        sc.calls = 0
        _active_counters[self.index] = sc
        return sc

    def __exit__(self, *exc):
        sc = _active_counters.pop(self.index)
        if exc[0] is None and sc.calls:
            with torch.cuda.device(self.index):
                check(load().smaq_counter_add(sc.base_ptr, sc.calls, stream_ptr(self.index)), "smaq_counter_add")
        return False


# Pinned staging for descriptor uploads made WHILE A GRAPH IS BEING CAPTURED: the captured host-to-device copy reads
# its source at every replay, and allocating pinned memory during capture is not allowed — so slices of one arena,
# allocated at the first (eager) batched call and never freed.
_ARENA_BYTES = 8 << 20
_arena = None
_arena_used = 0


def ensure_pinned_arena():
    global _arena
    if _arena is None and not torch.cuda.is_current_stream_capturing():
        _arena = torch.empty(_ARENA_BYTES, dtype=torch.uint8).pin_memory()


def pinned_arena_take(payload: bytes) -> torch.Tensor:
    global _arena_used
    if _arena is None:
        raise NativeLibraryError("a batched codec call must run once eagerly before it is captured in a CUDA graph "
                                 "(the pinned staging arena is allocated then)")
    n = len(payload)
    start = (_arena_used + 63) // 64 * 64
    if start + n > _ARENA_BYTES:
        raise NativeLibraryError("pinned staging arena exhausted (too many captured descriptor uploads)")
    _arena_used = start + n
    view = _arena[start:start + n]
    view.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    return view


def ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def is_dense(t: torch.Tensor) -> bool:
    """The tensor's elements fill one contiguous block of storage in SOME order (contiguous, channels_last, a permuted
    view of a contiguous tensor ...).  Statistics and the elementwise round trip do not depend on the order, so such
    tensors are processed in storage order and the result keeps their strides (``torch.empty_like`` preserves them) —
    like the reference's elementwise chain, and unlike ``.contiguous()``, which would add a transpose copy and hand
    NCHW tensors to a channels_last network."""
    return t.is_contiguous() or bool(torch.ops.aten.is_non_overlapping_and_dense(t))


def storage_order(t: torch.Tensor) -> torch.Tensor:
    """1-D view of a dense tensor's elements in storage order."""
    if t.is_contiguous():
        return t.view(-1)
    return t.as_strided((t.numel(),), (1,), t.storage_offset())


def require_cuda_f32(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise NativeLibraryError(
            f"{what}: tensor is on {t.device}; this implementation is CUDA (sm_100a) only and has no CPU path"
        )
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype}")
