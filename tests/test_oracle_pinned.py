"""Pin the CPU oracle (oracle/smaq.py) to the reference's own outputs.

tests/golden/smaq_reference_vectors.npz was produced by running the UNMODIFIED
reference SmartFP (smart_compress/compress/smart.py:110-190) in the build
container; these tests demand bit equality.  Where /root/reference is present
the live reference is also run on fresh random inputs.
"""
import pytest
import torch

from oracle.smaq import SmaqConfig, smaq_roundtrip
from tests.golden_util import assert_bit_equal, load_golden

CASES = load_golden()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    c = CASES[name]
    res = smaq_roundtrip(c["x"].clone(), c["cfg"], probs=c["probs"], idx=c["idx"], **c["kwargs"])
    assert res.passthrough == c["same_object"]
    assert res.y.dtype == c["y"].dtype and res.y.shape == c["y"].shape
    assert_bit_equal(res.y, c["y"], name)


def test_known_answer_vector():
    """SURVEY.md §4 KAT, generated from the reference with --no_stochastic_rounding."""
    c = CASES["kat_trunc"]
    res = smaq_roundtrip(c["x"], c["cfg"])
    assert res.mean.view(torch.int32).item() == 0x3F555555
    assert res.std.view(torch.int32).item() == 0x405774C3
    want = [0xC03BC5B6, 0xBFEE1028, 0xBFB49BB0, 0xBF764E6F, 0xBE93E20A, 0x3E23DFB0,
            0x3F1BE0DC, 0x3F555555, 0x3FA41F23, 0x3FFA4DD8, 0x40369B64, 0x411F8923]
    got = [v & 0xFFFFFFFF for v in res.y.view(torch.int32).tolist()]
    assert got == want


def test_default_ranges():
    cfg = SmaqConfig()
    assert cfg.range_normal == 15.0 and cfg.range_outlier == 42.0  # smart.py:75-80


@pytest.mark.container
@pytest.mark.parametrize("argv", [[], ["--no_stochastic_rounding"], ["--use_sample_stats"],
                                  ["--num_bits_main", "5", "--num_bits_outlier", "7"]])
def test_oracle_matches_live_reference(argv):
    from oracle import refload
    from tests.golden_util import parse_argv

    fp = refload.load_reference_smartfp(argv)
    cfg = parse_argv(" ".join(argv))
    g = torch.Generator().manual_seed(99)
    for n in (8, 33, 1 << 14, (1 << 18) + 3):
        x = torch.randn(n, generator=g)
        x[torch.randperm(n, generator=g)[: max(1, n // 100)]] *= 10
        probs = torch.rand(n, generator=g)
        perm = torch.randperm(n, generator=g)
        rl, rp = torch.rand_like, torch.randperm
        torch.rand_like = lambda t, **k: probs.clone()
        torch.randperm = lambda m, **k: perm.clone()
        try:
            y_ref = fp(x.clone(), tag="t")
        finally:
            torch.rand_like, torch.randperm = rl, rp
        res = smaq_roundtrip(x, cfg, probs=probs, idx=perm[: min(n, cfg.num_samples)])
        assert_bit_equal(res.y, y_ref, f"n={n} argv={argv}")
