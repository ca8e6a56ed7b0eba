"""Generate tests/golden/smaq_reference_vectors.npz by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

Each case feeds a seeded input to the reference's own ``SmartFP`` (imported
from /root/reference through oracle/refload.py) with the reference's own flag
parser.  ``torch.rand_like`` / ``torch.randperm`` are replaced for the duration
of the call by functions that return arrays we also store, so that the oracle
and the CUDA kernels can be driven with the very same randomness.  Only the
reference's *outputs* are trusted; nothing is computed by our code here.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refload  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_input(kind: str, n: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    if kind == "normal":
        return torch.randn(n, generator=g)
    if kind == "outliers":  # BASELINE.md §4 input recipe
        x = torch.randn(n, generator=g)
        k = max(1, n // 100)
        idx = torch.randperm(n, generator=g)[:k]
        x[idx] *= 10
        return x
    if kind == "shifted":
        return torch.randn(n, generator=g) * 0.02 + 3.0
    if kind == "positive":
        return torch.randn(n, generator=g).square() * 1e-3
    if kind == "constant":
        return torch.full((n,), 0.75)
    if kind == "kat":
        return torch.tensor([-3, -2, -1.5, -1, -0.5, 0, 0.5, 1, 1.5, 2, 3, 10], dtype=torch.float32)
    if kind == "nan":
        x = torch.randn(n, generator=g)
        x[n // 2] = float("nan")
        return x
    if kind == "tiny":
        return torch.randn(n, generator=g) * 1e-30
    if kind == "image":  # a 4-D feature map, as the autograd hook sees it
        return torch.randn(n, generator=g).view(2, -1, 4, 4).relu()
    if kind == "image3":  # batch of 3 (N != W != C, so a wrong permutation cannot pass by accident)
        return (torch.randn(n, generator=g) + 0.3).view(3, -1, 2, 8)
    raise ValueError(kind)


# name, input kind, n, seed, reference argv, call kwargs, precision
CASES = [
    ("kat_trunc", "kat", 12, 0, ["--no_stochastic_rounding"], {}, 32),
    ("n8_trunc", "normal", 8, 1, ["--no_stochastic_rounding"], {}, 32),
    ("n7_passthrough", "normal", 7, 2, [], {}, 32),
    ("n9_sr", "normal", 9, 3, [], {}, 32),
    ("n1000_sr", "outliers", 1000, 4, [], {}, 32),
    ("n4099_sr", "outliers", 4099, 5, [], {}, 32),
    ("n4099_trunc", "outliers", 4099, 6, ["--no_stochastic_rounding"], {}, 32),
    ("n65536_sr", "outliers", 65536, 7, [], {}, 32),
    ("shifted_sr", "shifted", 5000, 8, [], {}, 32),
    ("allpos_sr", "positive", 3001, 9, [], {"all_positive": True}, 32),
    ("constant_sr", "constant", 257, 10, [], {}, 32),
    ("nan_sr", "nan", 100, 11, [], {}, 32),
    ("tiny_sr", "tiny", 513, 12, [], {}, 32),
    ("sampled_sr", "outliers", 4099, 13, ["--use_sample_stats"], {}, 32),
    ("sampled64_trunc", "normal", 2050, 14, ["--use_sample_stats", "--num_samples", "64", "--no_stochastic_rounding"], {}, 32),
    ("sampled_small", "normal", 11, 15, ["--use_sample_stats"], {}, 32),
    ("bits_4_6_sr", "outliers", 2000, 16, ["--num_bits_main", "4", "--num_bits_outlier", "6"], {}, 32),
    ("bits_5_9_thr", "outliers", 2000, 17,
     ["--num_bits_main", "5", "--num_bits_outlier", "9", "--main_std_dev_threshold", "1.1",
      "--outlier_std_dev_threshold", "2.7"], {}, 32),
    ("range_std_sr", "normal", 3000, 18, ["--use_range_std_dev"], {}, 32),
    ("prec16_clamp", "tiny", 600, 19, [], {}, 16),
    ("image_sr", "image", 2 * 5 * 16, 20, [], {}, 32),
    ("sampled_range_sr", "outliers", 3000, 26, ["--use_sample_stats", "--use_range_std_dev"], {}, 32),
    # --use_batch_norm (smart.py:121,136-149,174-179): an NCHW feature map with the layer's gamma / beta; the
    # marker "bn" is replaced by seeded (gamma, beta) of C entries, stored next to the case
    ("bn_sr", "image", 2 * 5 * 16, 21, ["--use_batch_norm"], {"batch_norm_stats": "bn"}, 32),
    ("bn_scalar_sr", "image", 2 * 5 * 16, 22, ["--use_batch_norm", "--bn_scalar_params"], {"batch_norm_stats": "bn"}, 32),
    ("bn_allpos_trunc", "image3", 3 * 7 * 16, 23, ["--use_batch_norm", "--no_stochastic_rounding"],
     {"batch_norm_stats": "bn", "all_positive": True}, 32),
    ("bn_flag_without_stats_sr", "image", 2 * 5 * 16, 24, ["--use_batch_norm"], {}, 32),
    ("bn_stats_without_flag_sr", "image", 2 * 5 * 16, 25, [], {"batch_norm_stats": "bn"}, 32),
]


def run_case(name, kind, n, seed, argv, kwargs, precision):
    x = make_input(kind, n, seed)
    g = torch.Generator().manual_seed(1000 + seed)
    probs = torch.rand(x.shape, generator=g)
    perm = torch.randperm(x.numel(), generator=g)
    fp = refload.load_reference_smartfp(argv, precision=precision)
    kwargs = dict(kwargs)
    bn = None
    if kwargs.get("batch_norm_stats") == "bn":  # per-channel affine parameters of the producing BatchNorm2d
        channels = x.shape[1]
        bn = (torch.rand(channels, generator=g) + 0.5, torch.randn(channels, generator=g) * 0.1)
        kwargs["batch_norm_stats"] = bn

    real_rand_like, real_randperm = torch.rand_like, torch.randperm
    used = {"probs": False, "perm": False}

    def fake_rand_like(t, **kw):
        assert t.shape == probs.shape
        used["probs"] = True
        return probs.clone()

    def fake_randperm(m, **kw):
        assert m == perm.numel()
        used["perm"] = True
        return perm.clone()

    torch.rand_like, torch.randperm = fake_rand_like, fake_randperm
    try:
        xin = x.clone()
        y = fp(xin, tag="golden", **kwargs)
    finally:
        torch.rand_like, torch.randperm = real_rand_like, real_randperm

    k = min(x.numel(), fp.hparams.num_samples)
    out = {
        f"{name}/x": x.numpy(),
        f"{name}/y": y.numpy(),
        f"{name}/argv": np.array(" ".join(argv)),
        f"{name}/kwargs": np.array(repr({k: v for k, v in kwargs.items() if k != "batch_norm_stats"})),
        f"{name}/precision": np.array(precision),
        f"{name}/same_object": np.array(y is xin),
    }
    if used["probs"]:
        out[f"{name}/probs"] = probs.numpy()
    if used["perm"]:
        out[f"{name}/idx"] = perm[:k].numpy()
    if bn is not None:
        out[f"{name}/gamma"], out[f"{name}/beta"] = bn[0].numpy(), bn[1].numpy()
    return out


def main():
    if not refload.reference_available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    blob = {"torch_version": np.array(torch.__version__)}
    for case in CASES:
        blob.update(run_case(*case))
    path = os.path.join(OUT, "smaq_reference_vectors.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes,", len(CASES), "cases")


if __name__ == "__main__":
    main()
