"""The flag surface (``smart_compress/util/train.py``) against the reference's own parser.

Container test: the reference's ``init_model_from_args`` (util/train.py:74-184) runs unmodified up to the end of
its two-phase parse — the stand-in Trainer of oracle/refload.py stops it there — and every key it produced that
belongs to the compression surface (phase-1 switches, the chosen codec's own flags, ``precision``, the ``*_name``
mirrors) must come out of ``parse_compression_args`` with the same value for the same argv."""
import pytest

from smart_compress.util.train import (DATA_STRUCTURES, build_compression, compression_argv, parse_compression_args)

ARGVS = [
    [],
    ["--compress", "smart"],
    ["--compress", "smart", "--num_bits_main", "5", "--num_bits_outlier", "9", "--no_compress_weights", "--compress_loss"],
    ["--compress", "smart", "--use_sample_stats", "--num_samples", "64", "--no_stochastic_rounding",
     "--main_std_dev_threshold", "1.25", "--outlier_std_dev_threshold", "3", "--min_size", "16",
     "--use_range_std_dev", "--use_batch_norm", "--bn_scalar_params", "--measure_compression_ratio"],
    ["--compress", "fp8", "--no_compress_forward", "--no_compress_backward", "--no_float_quantize_check_inf"],
    ["--compress", "s2fp8", "--no_compress_gradients", "--no_compress_momentum_vectors", "--precision", "16"],
    ["--compress", "bf16", "--no_compress"],
    ["--compress", "fp16", "--compression_hook_fn", "global_hook"],
    ["--compress", "fp32", "--compression_hook_fn", "autograd"],
    # the reference's own launch line (scripts/train.ps1:1) minus the Trainer's flags
    ["--model", "resnet", "--dataset", "cifar10", "--compress", "smart", "--batch_size", "128"],
]

SURFACE_PREFIXES = ("compress", "compression", "num_samples", "num_bits", "use_", "bn_", "main_std", "outlier_std", "min_size",
                    "stochastic_rounding", "measure_compression_ratio", "float_quantize", "precision")


def surface(flat):
    return {k: v for k, v in flat.items() if k.startswith(SURFACE_PREFIXES)}


def flatten(ns):
    return {k: (getattr(v, "__name__", v) if (isinstance(v, type) or callable(v)) else v) for k, v in vars(ns).items()}


@pytest.mark.container
@pytest.mark.parametrize("argv", ARGVS, ids=lambda a: " ".join(a) or "defaults")
def test_same_namespace_as_the_reference_parser(argv):
    from oracle import refload

    _, ref = refload.reference_parse_args(argv)
    ours = flatten(parse_compression_args(argv))
    ref_s, our_s = surface(ref), surface(ours)
    assert set(ref_s) == set(our_s), set(ref_s) ^ set(our_s)
    for k in sorted(ref_s):
        assert ref_s[k] == our_s[k], (k, ref_s[k], our_s[k])


@pytest.mark.container
@pytest.mark.parametrize("bad", [["--compress", "int8"], ["--compression_hook_fn", "nope"],
                                 ["--compress", "smart", "--num_bits_main"], ["--compress", "fp8", "--num_bits_main", "5"]])
def test_rejects_what_the_reference_rejects(bad):
    from oracle import refload

    with pytest.raises(SystemExit):
        refload.reference_parse_args(bad)
    with pytest.raises(SystemExit):
        parse_compression_args(bad, strict=True)   # the reference's phase 2 is a strict parse (train.py:184)


def test_defaults_without_the_reference():
    a = parse_compression_args([])
    assert a.compress and a.compression_cls.__name__ == "FP32" and a.compression_hook_fn.__name__ == "register_autograd_module"
    assert all(getattr(a, f"compress_{w}") for w in DATA_STRUCTURES) and a.compress_loss is False and a.precision == 32
    assert a.compression_cls_name == "smart_compress.compress.fp32.FP32"
    b = parse_compression_args("--compress smart --no_compress_momentum_vectors")  # a string is split (train.py:91-92)
    assert b.compression_cls.__name__ == "SmartFP" and not b.compress_momentum_vectors and b.num_bits_outlier == 8


def test_build_compression_wires_like_the_reference():
    import torch
    import torch.nn as nn

    from smart_compress.util.globals import Globals
    from smart_compress.util.pytorch.optimizer import OptimLP

    net = nn.Sequential(nn.Linear(4, 4), nn.Tanh())
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    args = parse_compression_args(compression_argv("fp32", only=["forward", "weights"]))
    codec, model, wrapped = build_compression(args, net, opt, log=lambda *a, **k: None)
    assert type(codec).__name__ == "FP32" and Globals.compression is codec and codec.log is not None
    assert isinstance(wrapped, OptimLP) and wrapped.weight_quant is not None and wrapped.grad_quant is None
    assert "new_forward" in model[0].forward.__name__          # the layer's forward was re-bound
    args = parse_compression_args(["--compress", "fp32", "--no_compress"])
    net2 = nn.Sequential(nn.Linear(4, 4))
    codec, model, wrapped = build_compression(args, net2, opt)
    assert codec is None and wrapped is opt and "new_forward" not in getattr(model[0].forward, "__name__", "")
