"""The boundary above the codec on a real device: autograd hooks and the optimizer wrapper
(reference smart_compress/util/pytorch/{autograd,hooks,optimizer}.py)."""
from argparse import ArgumentParser, Namespace
from collections import Counter

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def hparams(**kw):
    from smart_compress.compress.smart import SmartFP

    args = SmartFP.add_argparse_args(ArgumentParser()).parse_args([])
    args.precision = 32
    for k in ("compress_forward", "compress_backward", "compress_weights", "compress_gradients",
              "compress_momentum_vectors"):
        setattr(args, k, True)
    for k, v in kw.items():
        setattr(args, k, v)
    return args


class Recorder:
    """Wraps a plugin and records (tag, numel, kwargs) of each call."""

    def __init__(self, inner):
        self.inner, self.calls = inner, []

    def __call__(self, t, tag=None, **kw):
        self.calls.append((tag, t.numel(), {k: v for k, v in kw.items() if k == "all_positive"}))
        return self.inner(t, tag=tag, **kw)


def small_net():
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(), nn.AdaptiveAvgPool2d(1),
                         nn.Flatten(), nn.Linear(8, 10)).to(DEV)


@pytest.mark.parametrize("optim_name", ["sgd", "adamw"])
def test_training_step_through_hooks(optim_name):
    from smart_compress.compress.smart import SmartFP
    from smart_compress.util.pytorch.autograd import register_autograd_module
    from smart_compress.util.pytorch.hooks import wrap_optimizer

    torch.manual_seed(0)
    hp = hparams()
    codec = Recorder(SmartFP(hp))
    model = register_autograd_module(small_net(), codec, hp)
    bn = [p for m in model.modules() if type(m) == nn.BatchNorm2d for p in m.parameters(recurse=False)]
    rest = [p for m in model.modules() if type(m) != nn.BatchNorm2d for p in m.parameters(recurse=False)]
    groups = [dict(params=bn, no_weight_compression=True), dict(params=rest)]
    inner = (torch.optim.SGD(groups, lr=0.1, momentum=0.9) if optim_name == "sgd"
             else torch.optim.AdamW(groups, lr=1e-3))
    opt = wrap_optimizer(inner, codec, hp)
    x = torch.randn(16, 3, 8, 8, device=DEV)
    target = torch.randint(0, 10, (16,), device=DEV)
    losses = []
    for _ in range(3):
        def closure():
            opt.zero_grad()
            loss = nn.functional.cross_entropy(model(x), target)
            loss.backward()
            return loss

        losses.append(float(opt.step(closure)))
    assert all(torch.isfinite(torch.tensor(losses)))
    tags = Counter(t for t, _, _ in codec.calls)
    n_params = len(bn) + len(rest)
    # wrapped: Sequential itself, Conv2d, BatchNorm2d, ReLU, AdaptiveAvgPool2d, Linear (Flatten is not
    # in the predicate, quantization.py:166-184); every wrapped output needs a gradient
    assert tags["forward_autograd"] == 3 * 6
    assert tags["backward_autograd"] == 3 * 6
    assert tags["optimizer_grad"] == 3 * 2 * n_params           # before and after the update
    assert tags["optimizer_weight"] == 3 * len(rest)            # BN group opts out
    per_param_state = 1 if optim_name == "sgd" else 2
    assert tags["optimizer_momentum"] == 3 * per_param_state * n_params
    if optim_name == "adamw":
        assert sum(1 for t, _, kw in codec.calls if kw.get("all_positive")) == 3 * n_params
        for p in bn + rest:
            assert bool((inner.state[p]["exp_avg_sq"] >= 0).all())


def test_backward_runs_on_autograd_thread_and_matches_forward_stream():
    """Gradient maps are compressed on autograd's worker thread; results must be finite and the
    hook must not touch leaves that need no grad."""
    from smart_compress.compress.smart import SmartFP
    from smart_compress.util.pytorch.autograd import Compressor

    hp = hparams(stochastic_rounding=False)
    comp = Compressor(SmartFP(hp))
    x = torch.randn(64, 64, device=DEV, requires_grad=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        y = comp(x * 2.0)
        y.sum().backward()
    s.synchronize()
    assert x.grad is not None and bool(torch.isfinite(x.grad).all())
    leaf = torch.randn(64, device=DEV)  # no grad needed: backward returns (None, None)
    out = comp(leaf)
    assert not out.requires_grad


def test_compress_many_matches_the_oracle_given_its_statistics():
    """The batched optimizer-side path (smaq_roundtrip_multi: what every gradient, weight and state tensor of
    configs 3-5 goes through) against the ORACLE: per tensor, the statistics it reports are within 1e-6 of the
    reference's mean / unbiased std, and every output element equals the oracle's round trip given those statistics
    and the uniforms the kernel drew — bit for bit, for one-block tensors and for the ones cut into work items."""
    from smart_compress.compress.smart import SmartFP
    from oracle import rng as orng
    from oracle.smaq import SmaqConfig, full_mean_std, smaq_roundtrip

    g = torch.Generator().manual_seed(5)
    sizes = [10, 64, 512, 513, 4096, 32768, 7, 40000, 2048, 1 << 20, (1 << 18) + 3, 32769]   # 7 < min_size
    tensors = [torch.randn(n, generator=g) * (0.5 + i) + 0.1 * i for i, n in enumerate(sizes)]
    kwargs = [dict(all_positive=(i % 3 == 0)) for i in range(len(sizes))]
    torch.manual_seed(11)
    b = SmartFP(hparams())
    mine = [t.to(DEV) for t in tensors]
    stats = {}
    got = b.compress_many(mine, kwargs, tag="optimizer_momentum", stats_out=stats)
    cfg = SmaqConfig()
    stream = 0
    for i, n in enumerate(sizes):
        if n < 8:
            assert got[i] is mine[i] and torch.equal(got[i].cpu(), tensors[i]) and i not in stats
            continue
        ms = stats[i].cpu()
        ref_mean, ref_std = full_mean_std(tensors[i], cfg)
        assert abs(ms[0].item() - ref_mean.item()) <= 1e-6 * max(abs(ref_mean.item()), ref_std.item()), (i, n)
        assert abs(ms[1].item() - ref_std.item()) <= 1e-6 * ref_std.item(), (i, n)
        # the uniforms tensor i was rounded with: stream = first call number + its index among the quantised ones
        probs = torch.from_numpy(orng.probs_for(n, seed=11, offset=stream))   # oracle/rng.py: the kernels' generator
        stream += 1
        ref = smaq_roundtrip(tensors[i], cfg, probs=probs, mean=ms[0], std=ms[1], rng_rule=True, **kwargs[i])
        assert torch.equal(got[i].cpu().view(torch.int32), ref.y.view(torch.int32)), f"tensor {i} ({n} elements)"
    # and the per-tensor entry point draws the same streams: one-block tensors are bit-identical to it
    torch.manual_seed(11)
    a = SmartFP(hparams())
    want = [a(t.to(DEV), tag="optimizer_momentum", **kw) for t, kw in zip(tensors, kwargs)]
    for i, n in enumerate(sizes):
        if 8 <= n <= 32768:
            assert torch.equal(want[i].view(torch.int32), got[i].view(torch.int32)), f"tensor {i} ({n} elements)"


def test_tensor_on_another_device_than_the_current_one():
    """the C entry points launch on the CURRENT device; a tensor on cuda:1 while cuda:0 is current
    (model.to("cuda:1") without set_device) must still be compressed on its own GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from smart_compress.compress.fp8 import FP8
    from smart_compress.compress.s2fp8 import S2FP8
    from smart_compress.compress.smart import SmartFP

    hp = hparams()
    hp.float_quantize_check_inf = True
    assert torch.cuda.current_device() == 0
    g = torch.Generator().manual_seed(2)
    for n in (100, 5000, 1 << 20):
        x = torch.randn(n, generator=g)
        x1 = x.to("cuda:1")
        torch.manual_seed(5)
        fp = SmartFP(hp)
        y1 = fp(x1, tag="t")
        assert y1.device == x1.device and torch.cuda.current_device() == 0
        torch.manual_seed(5)
        y0 = SmartFP(hp)(x.to("cuda:0"), tag="t")
        assert torch.equal(y0.cpu(), y1.cpu())
        packed = fp.encode(x1)
        assert fp.decode(packed).device == x1.device
        for cls in (FP8, S2FP8):
            z1 = cls(hp)(x1, tag="t")
            assert z1.device == x1.device and bool(torch.isfinite(z1).all())
    many = [torch.randn(300, generator=g).to("cuda:1"), torch.randn(70000, generator=g).to("cuda:0"),
            torch.randn(64, generator=g).to("cuda:1")]
    out = SmartFP(hp).compress_many([t.clone() for t in many], None, tag="optimizer_grad")
    for t, o in zip(many, out):
        assert o.device == t.device and not torch.equal(o, t) and float((o - t).abs().max()) < 1.0
    torch.cuda.synchronize("cuda:1")


def test_wrapped_optimizer_uses_the_batched_path():
    from smart_compress.compress.smart import SmartFP
    from smart_compress.util.pytorch.hooks import wrap_optimizer

    hp = hparams()
    codec = SmartFP(hp)
    seen = []
    orig = codec.compress_many
    codec.compress_many = lambda tensors, kwargs_list=None, tag=None: (seen.append((tag, len(tensors))),
                                                                      orig(tensors, kwargs_list, tag=tag))[1]
    model = small_net()
    opt = wrap_optimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9), codec, hp)
    x = torch.randn(8, 3, 8, 8, device=DEV)

    def closure():
        opt.zero_grad()
        loss = model(x).square().mean()
        loss.backward()
        return loss

    opt.step(closure)
    n = len(list(model.parameters()))
    assert seen == [("optimizer_grad", n), ("optimizer_grad", n), ("optimizer_weight", n), ("optimizer_momentum", n)]
    for p in model.parameters():
        assert bool(torch.isfinite(p).all())


def test_packed_saved_tensors_cut_activation_memory_and_keep_gradients_close():
    from smart_compress.compress.smart import SmartFP
    from smart_compress.util.pytorch.autograd import packed_saved_tensors

    torch.manual_seed(3)
    codec = SmartFP(hparams())
    # a ReLU network: the saved ReLU outputs must keep their exact zeros (the backward mask is `output > 0`), which
    # the encoder's zero_on_grid option guarantees
    net = nn.Sequential(nn.Conv2d(8, 32, 3, padding=1), nn.ReLU(), nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(),
                        nn.Conv2d(32, 8, 3, padding=1)).to(DEV)
    x = torch.randn(16, 8, 64, 64, device=DEV)

    def run(packed):
        net.zero_grad()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        if packed:
            with packed_saved_tensors(codec, min_numel=1 << 12):
                loss = net(x).square().mean()
        else:
            loss = net(x).square().mean()
        held = torch.cuda.memory_allocated() - base   # what the graph keeps alive for backward
        loss.backward()
        return held, [p.grad.clone() for p in net.parameters()]

    held_plain, g_plain = run(False)
    held_packed, g_packed = run(True)
    # saved activations shrink ~5x (exact-size streams: 6.4 bits per element); the network output and the packed
    # copy of the (externally owned) input stay, so the whole graph here shrinks about 2x
    assert held_packed < 0.6 * held_plain, (held_packed, held_plain)
    for a, b in zip(g_plain, g_packed):
        assert bool(torch.isfinite(b).all())
        rel = float((a - b).norm() / a.norm())
        assert rel < 0.15, rel   # 6/8-bit activations: a few percent of gradient noise
    # the same comparison WITHOUT zero_on_grid is what round 1 shipped: half of the dead units leak gradient
    net.zero_grad()
    with packed_saved_tensors(codec, min_numel=1 << 12, zero_on_grid=False):
        loss = net(x).square().mean()
    loss.backward()
    leaky = max(float((a - p.grad).norm() / a.norm()) for a, p in zip(g_plain, net.parameters()))
    tight = max(float((a - b).norm() / a.norm()) for a, b in zip(g_plain, g_packed))
    assert tight < leaky, (tight, leaky)


def _graph_train(codec_name, use_graph, steps=4):
    """`steps` SGD steps of a small network with the codec on all five data structures, every step a counted step;
    steps 2.. either eagerly or as replays of ONE captured CUDA graph."""
    from smart_compress import _native as N
    from smart_compress.util.train import build_compression, parse_compression_args

    torch.manual_seed(21)
    for sc in N._step_counters.values():
        sc.base.zero_()
    hp = parse_compression_args(["--compress", codec_name])
    net = nn.Sequential(nn.Linear(96, 256), nn.Tanh(), nn.LayerNorm(256), nn.Linear(256, 300), nn.Tanh(),
                        nn.Linear(300, 8)).to(DEV)
    x = torch.randn(64, 96, device=DEV)
    y = torch.randn(64, 8, device=DEV)
    inner = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
    codec, net, opt = build_compression(hp, net, inner)

    def closure():
        opt.zero_grad(set_to_none=True)
        loss = nn.functional.mse_loss(net(x), y)
        loss.backward()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with N.counted_step(DEV):
            opt.step(closure)                      # step 1, eager: lazy optimizer state, workspaces, plans
        side.synchronize()
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                with N.counted_step(DEV):
                    opt.step(closure)              # capture only: nothing runs
            for _ in range(steps - 1):
                graph.replay()
        else:
            for _ in range(steps - 1):
                with N.counted_step(DEV):
                    opt.step(closure)
        side.synchronize()
    torch.cuda.current_stream().wait_stream(side)
    counter = int(N._step_counters[torch.device(DEV).index].base.item())
    return torch.cat([p.detach().flatten() for p in net.parameters()]).clone(), counter


@pytest.mark.parametrize("codec_name", ["smart", "fp8"])
def test_cuda_graph_replay_of_a_whole_training_step_matches_eager_bits(codec_name):
    """§8 f-1: forward, backward, the hooks and the batched optimizer phases captured as ONE graph.  The random
    streams are numbered from the start of the step plus a DEVICE counter the step advances, so replays draw fresh
    numbers — the same ones eager execution draws: after four steps the weights are bit-identical, and so is the
    counter.  Two replays of the same graph must differ from one another (not the baked-offset bug)."""
    w_eager, c_eager = _graph_train(codec_name, use_graph=False)
    w_graph, c_graph = _graph_train(codec_name, use_graph=True)
    assert c_eager == c_graph and c_eager > 0
    assert torch.equal(w_eager.view(torch.int32), w_graph.view(torch.int32))
    w3, _ = _graph_train(codec_name, use_graph=True, steps=3)
    assert not torch.equal(w3, w_graph)


def test_counted_step_numbers_streams_from_a_device_counter():
    """The kernels add *offset_base to the stream offset: a call with (offset 3, counter 5) equals a call with
    (offset 8, no counter), for the round trip, the one-block kernel, the packed encoder and float_quantize."""
    from oracle.smaq import SmaqConfig
    from tests import cabi, cabi_pack

    g = torch.Generator().manual_seed(1)
    cfg = SmaqConfig()
    counter = torch.tensor([5], dtype=torch.int64, device=DEV)
    for n in (1000, 70001):
        x = torch.randn(n, generator=g).to(DEV)
        ms = cabi.stats_full(x)
        a, b = cabi.codec_params(cfg, seed=9, offset=3), cabi.codec_params(cfg, seed=9, offset=8)
        a.offset_base = counter.data_ptr()
        assert torch.equal(cabi.roundtrip(x, ms, a), cabi.roundtrip(x, ms, b))
        if n <= 32768:
            assert torch.equal(cabi.roundtrip_small(x, a), cabi.roundtrip_small(x, b))
        assert torch.equal(cabi_pack.encode(x, ms, a, cfg)[0], cabi_pack.encode(x, ms, b, cfg)[0])
        fa, fb = cabi.floatq_params(5, 2, seed=9, offset=3), cabi.floatq_params(5, 2, seed=9, offset=8)
        fa.offset_base = counter.data_ptr()
        assert torch.equal(cabi.float_quantize(x, fa), cabi.float_quantize(x, fb))
        assert not torch.equal(cabi.roundtrip(x, ms, a), cabi.roundtrip(x, ms, cabi.codec_params(cfg, seed=9, offset=3)))


def test_packed_saved_tensors_exact_size_delivers_the_accounted_ratio():
    """§8 f-2: the streams autograd holds are compacted to their used words without a synchronisation (the header is
    copied to pinned memory asynchronously, the compaction runs at a later pack / unpack call): what stays allocated
    per saved tensor is >= 4.8x smaller than fp32 (capacity-sized buffers: 4.0x), and compaction changes no value —
    the gradients are bit-identical to the capacity-sized run under the same seed."""
    from smart_compress.compress.smart import SmartFP
    from smart_compress.util.pytorch.autograd import packed_saved_tensors

    net = nn.Sequential(*[m for _ in range(6) for m in (nn.Conv2d(16, 16, 3, padding=1), nn.ReLU())]).to(DEV)
    x = torch.randn(8, 16, 96, 96, device=DEV)

    def run(exact):
        torch.manual_seed(5)
        codec = SmartFP(hparams())
        kept = []
        orig = codec.encode
        codec.encode = lambda t, **kw: (kept.append(orig(t, **kw)), kept[-1])[1]
        net.zero_grad()
        ctx = packed_saved_tensors(codec, min_numel=1 << 12, exact_size=exact)
        with ctx:
            loss = net(x).square().mean()
        torch.cuda.synchronize()
        with ctx:            # one more pack call after the copies have landed drains the queue
            torch.zeros(1 << 13, device=DEV, requires_grad=True).relu().sum()
        sizes = [(p.allocated_bytes(), 4 * p.numel) for p in kept[:-1]]
        loss.backward()
        return sizes, [p.grad.clone() for p in net.parameters()], ctx

    sizes_cap, g_cap, _ = run(False)
    sizes_exact, g_exact, ctx = run(True)
    assert ctx.compacted >= len(sizes_exact) - 1
    ratio_cap = sum(b for _, b in sizes_cap) / sum(a for a, _ in sizes_cap)
    ratio_exact = sum(b for _, b in sizes_exact) / sum(a for a, _ in sizes_exact)
    assert 3.9 < ratio_cap < 4.05 and ratio_exact >= 4.8, (ratio_cap, ratio_exact)
    for a, b in zip(g_cap, g_exact):
        assert torch.equal(a, b)
