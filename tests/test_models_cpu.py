"""The benchmark-driver network (smart_compress/models/pytorch/resnet.py) must expose the reference's hook
surface: the layer predicate wraps its blocks because of their module path, so the codec sees the reference's
call counts (76 forward calls per step for ResNet-18, 132 for ResNet-34).  CPU only."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smart-quantization_b200"))


@pytest.mark.parametrize("name,wrapped,calls,tensors,params", [("resnet18", 68, 76, 62, 11173962),
                                                                ("resnet34", 116, 132, 110, 21282122)])
def test_resnet_hook_surface(name, wrapped, calls, tensors, params):
    from smart_compress.models.pytorch.resnet import build
    from smart_compress.util.pytorch.quantization import is_valid_layer_type

    model = build(name)
    mods = [m for m in model.modules() if is_valid_layer_type(m)]
    seen = []
    for m in mods:
        m.register_forward_hook(lambda mod, inp, out: seen.append(out.numel()))
    y = model(torch.randn(2, 3, 32, 32))
    assert y.shape == (2, 10)
    assert len(mods) == wrapped and len(seen) == calls
    assert len(list(model.parameters())) == tensors and sum(p.numel() for p in model.parameters()) == params
