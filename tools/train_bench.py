#!/usr/bin/env python
"""Training-level measurement (BASELINE.json configs 3-5): synthetic-input, random-init ResNet-18/34 or
BERT-base training steps with the codec hooked onto every data structure the reference compresses —
feature maps and gradient maps (util/pytorch/autograd.py), gradients, weights and optimizer state
(util/pytorch/optimizer.py) — wired exactly as reference smart_compress/util/train.py:197-213 and
models/base.py:137-163 wire them (``smart_compress.util.train.parse_compression_args`` /
``build_compression``; BatchNorm parameters in a ``no_weight_compression`` group; SGD lr 0.1 momentum 0.9 for the
ResNets, AdamW for BERT; ``--compress_loss`` honoured, models/base.py:114-115).

    python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress smart --steps 50 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_bench.py \
        --model resnet34 --batch 32 --image 224 --compress smart

One process per GPU; DDP (NCCL over NVLink) issues the only collective, the bucketed gradient all-reduce, BEFORE
the optimizer-side compression — the reference's layout (optimizer.py:135-141).  ``run_training`` is also what
``bench.py`` calls for its ``train`` block.  ``--codec reference-eager`` swaps the CUDA kernels for the CPU oracle's
port of smart.py evaluated by torch's own CUDA operators through the SAME hooks: what running the reference
unchanged on this GPU costs (a baseline leg; the only place this file touches ``oracle/``).
"""
import argparse
import contextlib
import json
import os
import sys
import time
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402

DATA_STRUCTURES = ("forward", "backward", "weights", "gradients", "momentum_vectors")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18", choices=["resnet18", "resnet34", "bert-base"])
    ap.add_argument("--batch", type=int, default=256, help="per GPU")
    ap.add_argument("--image", type=int, default=32)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--compress", default="smart", choices=["smart", "fp8", "s2fp8", "fp16", "bf16", "fp32"])
    ap.add_argument("--codec", default="b200", choices=["b200", "reference-eager"])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--lr", type=float, default=None,
                    help="learning rate (default: SGD 0.1 / AdamW 2e-5). With 8-bit codecs on EVERY tensor and random data "
                         "0.1 diverges within ~50 steps; a run whose loss is not finite times the codecs' NaN paths")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--no-batched-optimizer", action="store_true",
                    help="per-tensor optimizer-side calls (the reference's loop) instead of compress_many")
    ap.add_argument("--packed-activations", nargs="?", const="exact", default="", choices=["exact", "capacity"],
                    help="keep autograd's saved tensors as packed SmaQ streams (not in the reference; changes numerics); "
                         "exact: compacted to their used words (5.0x), capacity: 8 bits per element (4.0x)")
    ap.add_argument("--compress_loss", action="store_true")
    ap.add_argument("--find-nonfinite", action="store_true",
                    help="debugging: report the first codec call that turns a finite tensor into a non-finite one "
                         "(synchronises every call; per-tensor optimizer calls)")
    ap.add_argument("--compress-allreduce", default="", choices=["", "p2p", "nccl"],
                    help="gradient compression fused with the all-reduce (smart_compress/util/pytorch/allreduce.py; not in the "
                         "reference, changes numerics): DDP's fp32 all-reduce is replaced by packed SmaQ streams read over NVLink")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="capture the whole training step (forward, backward, hooks, optimizer) in ONE CUDA graph and replay "
                         "it: the codec's random streams advance through a device counter (smart_compress._native.counted_step). "
                         "Single GPU only")
    ap.add_argument("--profile", action="store_true",
                    help="after the timed steps, run 3 more under torch.profiler and print GPU-busy time and the top kernels to stderr")
    ap.add_argument("--only", default=",".join(DATA_STRUCTURES),
                    help="which data structures are compressed (reference --no_compress_* flags)")
    return ap.parse_args()


class ReferenceEagerSmaQ:
    """BASELINE LEG ONLY: the reference's SmartFP.__call__ (smart.py:110-190) as restated op for op by
    oracle/smaq.py, evaluated by torch's own CUDA operators — ~28 launches, rand_like and the blocking
    ``if std_dev == 0`` read per call — behind the same hook boundary as the product."""

    def __init__(self, hparams):
        from oracle.smaq import SmaqConfig, smaq_roundtrip

        self.hparams = hparams
        self.cfg = SmaqConfig()
        self._rt = smaq_roundtrip
        self.log = self.log_custom = None

    @torch.no_grad()
    def __call__(self, t, tag=None, all_positive=False, **_):
        if t.numel() < self.cfg.min_size:
            return t
        return self._rt(t, self.cfg, probs=torch.rand_like(t), all_positive=all_positive).y


def build_model(model_name, batch, image, seq, device):
    if model_name.startswith("resnet"):
        from smart_compress.models.pytorch.resnet import build

        model = build(model_name, num_classes=10).to(device)
        x = torch.randn(batch, 3, image, image, device=device)
        y = torch.randint(0, 10, (batch,), device=device)

        def loss_fn(m):
            return nn.functional.cross_entropy(m(x), y)

        return model, loss_fn, "img/s"
    from transformers import BertConfig, BertForSequenceClassification

    cfg = BertConfig(num_labels=1)  # bert-base-uncased shape; STS-B is a regression task
    model = BertForSequenceClassification(cfg).to(device)
    ids = torch.randint(0, cfg.vocab_size, (batch, seq), device=device)
    mask = torch.ones_like(ids)
    y = torch.rand(batch, device=device) * 5

    def loss_fn(m):
        return m(input_ids=ids, attention_mask=mask, labels=y).loss

    return model, loss_fn, "seq/s"


def run_training(model_name="resnet34", batch=32, image=224, seq=128, compress="smart", steps=50, warmup=10,
                 device=None, world=1, local=0, codec="b200", only=DATA_STRUCTURES, batched_optimizer=True,
                 packed_activations=False, compress_loss_flag=False, profile=False, seed=1234, clocks=None,
                 cuda_graph=False, compress_allreduce="", lr=None, find_nonfinite=False):
    """`steps` timed training steps after `warmup` untimed ones; returns a dict (device-timed, max over ranks)."""
    from smart_compress.util.pytorch.autograd import packed_saved_tensors
    from smart_compress.util.train import build_compression, compress_loss, compression_argv, parse_compression_args

    rank = int(os.environ.get("RANK", 0))
    torch.manual_seed(seed + rank)  # the reference seeds nothing: per-rank rounding streams
    argv = compression_argv(compress, only=list(only), extra=["--compress_loss"] if compress_loss_flag else [])
    if compress == "fp32":
        argv.append("--no_compress")
    hp = parse_compression_args(argv)
    calls = Counter()
    nonfinite = {}
    if find_nonfinite:
        batched_optimizer = False

    class Counting:  # counts calls per tag; forwards compress_many when allowed
        def __init__(self, inner):
            self.inner = inner
            if hasattr(inner, "compress_many") and batched_optimizer:
                self.compress_many = self._many

        def __call__(self, t, tag=None, **kw):
            calls[tag] += 1
            out = self.inner(t, tag=tag, **kw)
            if find_nonfinite and not nonfinite and bool(torch.isfinite(t).all()) and not bool(torch.isfinite(out).all()):
                a = t.abs()
                nonfinite.update(call=sum(calls.values()), tag=tag, numel=t.numel(), shape=list(t.shape),
                                 absmax=float(a.max()), absmin_nonzero=float(a[a > 0].min()) if bool((a > 0).any()) else 0.0,
                                 zero_fraction=float((a == 0).float().mean()),
                                 bad_out_fraction=float((~torch.isfinite(out)).float().mean()))
            return out

        def _many(self, tensors, kwargs_list=None, tag=None):
            calls[f"{tag} (batched)"] += len(tensors)
            return self.inner.compress_many(tensors, kwargs_list, tag=tag)

    model, loss_fn, unit = build_model(model_name, batch, image, seq, device)
    # reference models/base.py:137-150: BatchNorm2d parameters never have their weights compressed
    bn = [p for m in model.modules() if type(m) == nn.BatchNorm2d for p in m.parameters(recurse=False)]
    rest = [p for m in model.modules() if type(m) != nn.BatchNorm2d for p in m.parameters(recurse=False)]
    groups = [dict(params=bn, no_weight_compression=True), dict(params=rest)] if bn else [dict(params=rest)]
    if model_name.startswith("resnet"):
        inner = torch.optim.SGD(groups, lr=0.1 if lr is None else lr, momentum=0.9, weight_decay=0)
    else:
        inner = torch.optim.AdamW(groups, lr=2e-5 if lr is None else lr, capturable=bool(cuda_graph))

    if codec == "reference-eager":
        assert compress == "smart"
        hp.compression_cls = ReferenceEagerSmaQ
    # util/train.py:197-213 + models/base.py:152-157, with the call counter between the hooks and the codec
    real_cls = hp.compression_cls
    hp.compression_cls = lambda args: Counting(real_cls(args))
    fn, model, opt = build_compression(hp, model, inner)
    pack_codec = None
    if packed_activations:
        from smart_compress.compress.smart import SmartFP

        pack_codec = fn.inner if isinstance(getattr(fn, "inner", None), SmartFP) else SmartFP(parse_compression_args(["--compress", "smart"]))
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    packed_ctx = packed_saved_tensors(pack_codec, exact_size=packed_activations != "capacity") if packed_activations else None
    car = None
    if compress_allreduce and world > 1:
        from smart_compress.compress.smart import SmartFP
        from smart_compress.util.pytorch.allreduce import register_compressed_allreduce

        wire_codec = fn.inner if isinstance(getattr(fn, "inner", None), SmartFP) else SmartFP(parse_compression_args(["--compress", "smart"]))
        car = register_compressed_allreduce(net, wire_codec, transport=compress_allreduce)

    def closure():
        opt.zero_grad(set_to_none=True)
        with (packed_ctx if packed_activations else contextlib.nullcontext()):
            loss = loss_fn(net)
        compress_loss(loss, fn, hp)   # models/base.py:114-115
        loss.backward()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_fn = lambda: opt.step(closure)  # noqa: E731
    graph = None
    if cuda_graph:
        # The whole step as ONE graph (§8 f-1): ~600-1200 codec launches plus the network's own kernels replayed
        # without a line of Python.  Warm-up runs eagerly on the capture stream (lazy optimizer state, workspaces,
        # cuDNN / cuBLAS plans); every step — eager or replayed — is a counted step, so the random streams advance
        # on the device exactly as they would eagerly.
        assert world == 1, "--cuda-graph is a single-GPU option (DDP's reducer is not captured here)"
        from smart_compress import _native as N

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 3)):
                with N.counted_step(device):
                    opt.step(closure)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                with N.counted_step(device):
                    static_loss = opt.step(closure)
        torch.cuda.current_stream().wait_stream(side)

        def step_fn():
            graph.replay()
            return static_loss
    else:
        for _ in range(warmup):
            opt.step(closure)
    sync()
    per_step_calls = {str(k): v // ((max(warmup, 3) + 1) if cuda_graph else max(warmup, 1)) for k, v in
                      sorted(calls.items(), key=lambda kv: str(kv[0]))}
    calls.clear()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with (clocks if clocks is not None else contextlib.nullcontext()):
        t0 = time.perf_counter()
        e0.record()
        last = None
        for _ in range(steps):
            last = step_fn()
        e1.record()
        sync()
        wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / steps
    out = {
        "value": round(batch * world / (ms_per_step / 1e3), 1), "unit": unit, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms_per_step, 3), "wall_ms_per_step": round(1e3 * wall / steps, 3),
        "loss": float(last), "peak_memory_gib": round(torch.cuda.max_memory_allocated() / 2**30, 2),
        "workload": f"{model_name} random-init, synthetic batch {batch}/GPU" +
                    (f" {image}x{image}" if model_name.startswith("resnet") else f" seq {seq}") +
                    f", --compress {compress} on {','.join(only)}" + (" [reference eager torch-CUDA ops]" if codec == "reference-eager" else ""),
        "optimizer": type(inner).__name__,
        "codec_calls_per_step": per_step_calls if cuda_graph else
                                {str(k): v // steps for k, v in sorted(calls.items(), key=lambda kv: str(kv[0]))},
        "cuda_graph": bool(cuda_graph),
        "first_nonfinite": nonfinite or None,
        "compress_allreduce": compress_allreduce or None,
        "allreduce_stats": None if car is None else dict(car.stats),
    }
    if profile and not cuda_graph:  # on EVERY rank: the extra steps contain DDP's collectives
        prof = profile_steps(opt, closure, ms_per_step)
        if rank == 0:
            out["profile"] = prof
    del net, model, opt, inner, graph, step_fn
    torch.cuda.empty_cache()
    return out


def profile_steps(opt, closure, ms_per_step, n=3):
    """GPU-busy time, codec-kernel time, NCCL time and the top kernels of `n` more steps (torch.profiler / CUPTI)."""
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            opt.step(closure)
        torch.cuda.synchronize()
    def is_kernel(e):  # device-side records of host annotations ("Optimizer.step#...", "DistributedDataParallel.forward") are not kernels
        return (e.device_type == torch.autograd.DeviceType.CUDA and not getattr(e, "is_user_annotation", False)
                and not e.name.startswith(("Optimizer.", "DistributedDataParallel", "ProfilerStep", "autograd::")))

    ev = [e for e in prof.events() if is_kernel(e)]
    per = 1e3 * n
    busy = sum(e.device_time for e in ev) / per
    ours = sum(e.device_time for e in ev if "smaq" in e.name or "floatq" in e.name) / per
    nccl = sum(e.device_time for e in ev if "nccl" in e.name.lower()) / per
    agg, cnt = Counter(), Counter()
    for e in ev:
        agg[e.name[:70]] += e.device_time / per
        cnt[e.name[:70]] += 1
    return {"gpu_busy_ms_per_step": round(busy, 3), "codec_kernels_ms_per_step": round(ours, 3),
            "nccl_kernels_ms_per_step": round(nccl, 3), "step_ms": round(ms_per_step, 3), "gpu_ops_per_step": len(ev) // n,
            "top": [{"ms": round(v, 3), "count": cnt[k] // n, "kernel": k} for k, v in agg.most_common(14)]}


def main():
    a = parse()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    r = run_training(a.model, a.batch, a.image, a.seq, a.compress, a.steps, a.warmup, device, world, local,
                     codec=a.codec, only=tuple(a.only.split(",")), batched_optimizer=not a.no_batched_optimizer,
                     packed_activations=a.packed_activations, compress_loss_flag=a.compress_loss, profile=a.profile,
                     cuda_graph=a.cuda_graph, compress_allreduce=a.compress_allreduce, lr=a.lr,
                     find_nonfinite=a.find_nonfinite)
    if rank == 0:
        prof = r.pop("profile", None)
        line = {"metric": f"{a.model}_train_{r['unit'].replace('/', '_per_')}", **r, "higher_is_better": True,
                "scaling": "weak", "dtype": "f32", "data": "synthetic", "packed_activations": bool(a.packed_activations),
                "config": {"workload": r["workload"], "optimizer": r["optimizer"],
                           "batched_optimizer_side": not a.no_batched_optimizer}}
        print(json.dumps(line))
        if prof:
            print(f"[profile] GPU busy {prof['gpu_busy_ms_per_step']} ms/step of {prof['step_ms']}; codec kernels "
                  f"{prof['codec_kernels_ms_per_step']} ms/step; NCCL {prof['nccl_kernels_ms_per_step']} ms/step; "
                  f"{prof['gpu_ops_per_step']} GPU ops/step", file=sys.stderr)
            for t in prof["top"]:
                print(f"[profile] {t['ms']:8.3f} ms  x{t['count']:<4d} {t['kernel']}", file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
