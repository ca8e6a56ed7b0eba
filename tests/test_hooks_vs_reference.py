"""The hook boundary against the REFERENCE'S OWN hooks (build container only: needs /root/reference).

The reference's ``register_autograd_module`` / ``wrap_optimizer`` / ``OptimLP``
(smart_compress/util/pytorch/{autograd,hooks,optimizer}.py) run unmodified here with three import stubs
(oracle/refload.py).  Both sides train the same seeded network for two steps with a RECORDING compress_fn — no codec,
so this runs on CPU — and the sequences of ``(tag, numel, all_positive, has batch_norm_stats)`` the codec would see
must be identical, call for call: the order of compress calls is the contract of the boundary
(reference optimizer.py:69-143, autograd.py:23-42,57-77).  The compress_fn adds a tag-dependent perturbation, so
the final weights also prove that each side assigns the results to the same tensors."""
import os
import sys
from argparse import Namespace

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]

pytestmark = pytest.mark.container


def hparams(**kw):
    hp = Namespace(compress_forward=True, compress_backward=True, compress_weights=True, compress_gradients=True,
                   compress_momentum_vectors=True, use_batch_norm=False, compress_loss=False)
    for k, v in kw.items():
        setattr(hp, k, v)
    return hp


class Recorder:
    """Stands in for the codec: records what it is called with, returns a fresh, slightly changed tensor."""

    def __init__(self):
        self.calls = []

    def __call__(self, t, tag=None, all_positive=False, batch_norm_stats=None, **kw):
        assert not kw, kw
        self.calls.append((tag, t.numel(), bool(all_positive), batch_norm_stats is not None,
                           None if batch_norm_stats is None else int(batch_norm_stats[0].numel())))
        scale = {"forward_autograd": 1.001, "backward_autograd": 0.999, "optimizer_grad": 1.002,
                 "optimizer_weight": 0.998, "optimizer_momentum": 1.003}[tag]
        return t * scale


def param_groups(model):
    """reference models/base.py:137-150: BatchNorm2d parameters are never weight-compressed."""
    bn, rest = [], []
    for child in model.modules():
        (bn if type(child) == nn.BatchNorm2d else rest).extend(child.parameters(recurse=False))
    return [dict(params=bn, no_weight_compression=True), dict(params=rest)]


def train(model, make_opt, register, wrap, hp, rec, batch, steps=2):
    model = register(model, rec, hp)
    opt = wrap(make_opt(param_groups(model)), rec, hp)
    x, y, loss_fn = batch

    def closure():
        opt.zero_grad()
        loss = loss_fn(model(x), y)
        loss.backward()
        return loss

    for _ in range(steps):
        opt.step(closure)
    return torch.cat([p.detach().flatten() for p in model.parameters()])


def reference_side(build_model, make_opt, hp, batch):
    from oracle import refload

    with refload.reference_modules():
        from smart_compress.util.pytorch.autograd import register_autograd_module
        from smart_compress.util.pytorch.hooks import wrap_optimizer

        rec = Recorder()
        torch.manual_seed(7)
        weights = train(build_model(True), make_opt, register_autograd_module, wrap_optimizer, hp, rec, batch)
    return rec.calls, weights


def our_side(build_model, make_opt, hp, batch, batched):
    from smart_compress.util.pytorch.autograd import register_autograd_module
    from smart_compress.util.pytorch.hooks import wrap_optimizer

    rec = Recorder()
    if batched:  # the batched form OptimLP uses when the codec offers it (SmartFP.compress_many)
        def many(tensors, kwargs_list=None, tag=None):
            kwargs_list = kwargs_list or [{}] * len(tensors)
            return [rec(t, tag=tag, **kw) for t, kw in zip(tensors, kwargs_list)]

        rec.compress_many = many
    torch.manual_seed(7)
    weights = train(build_model(False), make_opt, register_autograd_module, wrap_optimizer, hp, rec, batch)
    return rec.calls, weights


def resnet18(reference: bool):
    if reference:  # the reference's own network (smart_compress/models/pytorch/resnet.py:13-303)
        from smart_compress.models.pytorch.resnet import resnet18 as ref_resnet18

        return ref_resnet18(num_classes=10)
    from smart_compress.models.pytorch.resnet import build

    return build("resnet18", num_classes=10)


def small_mlp(reference: bool):
    return nn.Sequential(nn.Linear(24, 32), nn.Tanh(), nn.LayerNorm(32), nn.Linear(32, 1))


@pytest.mark.parametrize("batched", [False, True])
@pytest.mark.parametrize("use_bn", [False, True])
def test_resnet18_sgd_call_sequence_matches_the_reference_hooks(use_bn, batched):
    g = torch.Generator().manual_seed(3)
    batch = (torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 10, (4,), generator=g),
             nn.functional.cross_entropy)
    hp = hparams(use_batch_norm=use_bn)
    sgd = lambda groups: torch.optim.SGD(groups, lr=0.1, momentum=0.9)  # noqa: E731
    ref_calls, ref_w = reference_side(resnet18, sgd, hp, batch)
    our_calls, our_w = our_side(resnet18, sgd, hp, batch, batched)
    # 76 fwd + 76 bwd + 124 grad + 22 weight + 62 momentum per step (the first step has no
    # momentum buffers before the update, so they appear from its post-closure phase on)
    assert len(ref_calls) == len(our_calls)
    for i, (a, b) in enumerate(zip(ref_calls, our_calls)):
        assert a == b, f"call {i}: reference {a} vs ours {b}"
    per_step = len(ref_calls) // 2
    assert per_step == 76 + 76 + 124 + 22 + 62
    assert sum(c[3] for c in ref_calls) == (2 * 20 if use_bn else 0)  # 20 BatchNorm2d layers pass gamma / beta
    assert torch.equal(ref_w, our_w), "results were assigned to different tensors"


@pytest.mark.parametrize("batched", [False, True])
@pytest.mark.parametrize("flags", [dict(), dict(compress_backward=False, compress_weights=False),
                                   dict(compress_forward=False, compress_gradients=False,
                                        compress_momentum_vectors=False)])
def test_mlp_adamw_call_sequence_matches_the_reference_hooks(flags, batched):
    g = torch.Generator().manual_seed(4)
    batch = (torch.randn(16, 24, generator=g), torch.randn(16, 1, generator=g), nn.functional.mse_loss)
    hp = hparams(**flags)
    adamw = lambda groups: torch.optim.AdamW(groups, lr=1e-3)  # noqa: E731
    ref_calls, ref_w = reference_side(small_mlp, adamw, hp, batch)
    our_calls, our_w = our_side(small_mlp, adamw, hp, batch, batched)
    assert ref_calls == our_calls
    if hp.compress_momentum_vectors:  # exp_avg_sq travels with all_positive=True (optimizer.py:52-59)
        assert sum(c[2] for c in ref_calls) == 2 * 6
    assert torch.equal(ref_w, our_w)
