"""Read tests/golden/smaq_reference_vectors.npz (made by oracle/gen_golden.py)."""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def parse_argv(argv: str):
    """Reference flag string -> oracle SmaqConfig kwargs (flag names: smart.py:17-69)."""
    from oracle.smaq import SmaqConfig

    toks = argv.split()
    kw = {}
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "--no_stochastic_rounding":
            kw["stochastic_rounding"] = False
        elif t == "--use_sample_stats":
            kw["use_sample_stats"] = True
        elif t == "--use_range_std_dev":
            kw["use_range_std_dev"] = True
        elif t == "--use_batch_norm":
            kw["use_batch_norm"] = True
        elif t == "--bn_scalar_params":
            kw["bn_scalar_params"] = True
        elif t in ("--num_samples", "--num_bits_main", "--num_bits_outlier", "--min_size"):
            kw[t[2:]] = int(toks[i + 1])
            i += 1
        elif t in ("--main_std_dev_threshold", "--outlier_std_dev_threshold"):
            kw[t[2:]] = float(toks[i + 1])
            i += 1
        else:
            raise ValueError(t)
        i += 1
    return SmaqConfig(**kw)


def load_golden():
    z = np.load(os.path.join(HERE, "golden", "smaq_reference_vectors.npz"))
    names = sorted({k.split("/")[0] for k in z.files if "/" in k})
    cases = {}
    for n in names:
        c = {"name": n}
        for f in ("x", "y", "probs", "idx"):
            key = f"{n}/{f}"
            c[f] = torch.from_numpy(z[key].copy()) if key in z.files else None
        c["argv"] = str(z[f"{n}/argv"])
        c["kwargs"] = ast.literal_eval(str(z[f"{n}/kwargs"]))
        if f"{n}/gamma" in z.files:  # --use_batch_norm cases: the BatchNorm2d affine parameters
            c["kwargs"]["batch_norm_stats"] = (torch.from_numpy(z[f"{n}/gamma"].copy()), torch.from_numpy(z[f"{n}/beta"].copy()))
        c["precision"] = int(z[f"{n}/precision"])
        c["same_object"] = bool(z[f"{n}/same_object"])
        cfg = parse_argv(c["argv"])
        cfg.precision = c["precision"]
        c["cfg"] = cfg
        cases[n] = c
    return cases


def uses_bn(case) -> bool:
    """The case exercises --use_batch_norm (flag AND the layer's gamma / beta present, smart.py:121)."""
    return bool(case["cfg"].use_batch_norm and case["kwargs"].get("batch_norm_stats") is not None)


def bits(t: torch.Tensor) -> torch.Tensor:
    """fp32 tensor -> int32 bit pattern (for bit-exact comparison incl. NaN payloads and -0)."""
    return t.contiguous().view(torch.int32)


def assert_bit_equal(a: torch.Tensor, b: torch.Tensor, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ai, bi = bits(a.cpu()), bits(b.cpu())
    # every NaN is "the same NaN": torch CPU and CUDA need not agree on the payload
    an, bn = torch.isnan(a.cpu()), torch.isnan(b.cpu())
    assert torch.equal(an, bn), f"{what}: NaN positions differ"
    diff = (ai != bi) & ~an
    if diff.any():
        i = int(diff.flatten().nonzero()[0])
        raise AssertionError(
            f"{what}: {int(diff.sum())} of {a.numel()} differ; first at {i}: "
            f"{a.flatten()[i].item()!r} ({ai.flatten()[i].item():#x}) vs "
            f"{b.flatten()[i].item()!r} ({bi.flatten()[i].item():#x})"
        )
