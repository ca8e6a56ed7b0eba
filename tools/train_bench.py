#!/usr/bin/env python
"""Training-level measurement (BASELINE.json configs 3-5): synthetic-input, random-init ResNet-18/34 or
BERT-base training steps with the codec hooked onto every data structure the reference compresses —
feature maps and gradient maps (util/pytorch/autograd.py), gradients, weights and optimizer state
(util/pytorch/optimizer.py) — exactly as reference smart_compress/util/train.py:197-213 and
models/base.py:137-163 wire them (BatchNorm parameters in a ``no_weight_compression`` group; SGD
lr 0.1 momentum 0.9 for the ResNets, AdamW for BERT).

    python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress smart --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_bench.py \
        --model resnet34 --batch 32 --image 224 --compress smart

One process per GPU; DDP (NCCL over NVLink) issues the only collective, the bucketed gradient
all-reduce, BEFORE the optimizer-side compression — the reference's layout (optimizer.py:135-141).
Prints one JSON line: img/s (or seq/s) summed over ranks, device-timed, max over ranks.
"""
import argparse
import json
import os
import sys
import time
from argparse import ArgumentParser
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18", choices=["resnet18", "resnet34", "bert-base"])
    ap.add_argument("--batch", type=int, default=256, help="per GPU")
    ap.add_argument("--image", type=int, default=32)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--compress", default="smart", choices=["smart", "fp8", "s2fp8", "fp16", "bf16", "fp32"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-batched-optimizer", action="store_true",
                    help="per-tensor optimizer-side calls (the reference's loop) instead of compress_many")
    ap.add_argument("--packed-activations", action="store_true",
                    help="keep autograd's saved tensors as packed SmaQ streams (not in the reference; changes numerics)")
    ap.add_argument("--profile", action="store_true",
                    help="after the timed steps, run 3 more under torch.profiler and print GPU-busy time and the top kernels to stderr")
    ap.add_argument("--only", default="forward,backward,weights,gradients,momentum_vectors",
                    help="which data structures are compressed (reference --no_compress_* flags)")
    return ap.parse_args()


def codec_and_hparams(name, only):
    from smart_compress.compress import ALGORITHMS

    cls = ALGORITHMS[name]
    hp = cls.add_argparse_args(ArgumentParser()).parse_args([])
    hp.precision = 32
    for k in ("forward", "backward", "weights", "gradients", "momentum_vectors"):
        setattr(hp, f"compress_{k}", k in only and name != "fp32")
    hp.compress_loss = False
    return cls(hp), hp


def build_model(a, device):
    if a.model.startswith("resnet"):
        from smart_compress.models.pytorch.resnet import build

        model = build(a.model, num_classes=10).to(device)
        x = torch.randn(a.batch, 3, a.image, a.image, device=device)
        y = torch.randint(0, 10, (a.batch,), device=device)

        def loss_fn(m):
            return nn.functional.cross_entropy(m(x), y)

        unit, per_step = "img/s", a.batch
    else:
        from transformers import BertConfig, BertForSequenceClassification

        cfg = BertConfig(num_labels=1)  # bert-base-uncased shape; STS-B is a regression task
        model = BertForSequenceClassification(cfg).to(device)
        ids = torch.randint(0, cfg.vocab_size, (a.batch, a.seq), device=device)
        mask = torch.ones_like(ids)
        y = torch.rand(a.batch, device=device) * 5

        def loss_fn(m):
            return m(input_ids=ids, attention_mask=mask, labels=y).loss

        unit, per_step = "seq/s", a.batch
    return model, loss_fn, unit, per_step


def main():
    a = parse()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(1234 + rank)  # the reference seeds nothing: per-rank rounding streams

    from smart_compress.util.pytorch.autograd import register_autograd_module
    from smart_compress.util.pytorch.hooks import wrap_optimizer

    codec, hp = codec_and_hparams(a.compress, set(a.only.split(",")))
    calls = Counter()

    class Counting:  # counts calls per tag; forwards compress_many when allowed
        def __init__(self, inner):
            self.inner = inner
            if hasattr(inner, "compress_many") and not a.no_batched_optimizer:
                self.compress_many = self._many

        def __call__(self, t, tag=None, **kw):
            calls[tag] += 1
            return self.inner(t, tag=tag, **kw)

        def _many(self, tensors, kwargs_list=None, tag=None):
            calls[f"{tag} (batched)"] += len(tensors)
            return self.inner.compress_many(tensors, kwargs_list, tag=tag)

    fn = Counting(codec)
    model, loss_fn, unit, per_step = build_model(a, device)
    if hp.compress_forward or hp.compress_backward:
        model = register_autograd_module(model, fn, hp)
    # reference models/base.py:137-150: BatchNorm2d parameters never have their weights compressed
    bn = [p for m in model.modules() if type(m) == nn.BatchNorm2d for p in m.parameters(recurse=False)]
    rest = [p for m in model.modules() if type(m) != nn.BatchNorm2d for p in m.parameters(recurse=False)]
    groups = [dict(params=bn, no_weight_compression=True), dict(params=rest)] if bn else [dict(params=rest)]
    if a.model.startswith("resnet"):
        inner = torch.optim.SGD(groups, lr=0.1, momentum=0.9, weight_decay=0)
    else:
        inner = torch.optim.AdamW(groups, lr=2e-5)
    opt = wrap_optimizer(inner, fn, hp) if a.compress != "fp32" else inner
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model

    import contextlib

    from smart_compress.util.pytorch.autograd import packed_saved_tensors

    pack_codec = codec if hasattr(codec, "encode") else codec_and_hparams("smart", set())[0]

    def closure():
        opt.zero_grad(set_to_none=True)
        ctx = packed_saved_tensors(pack_codec) if a.packed_activations else contextlib.nullcontext()
        with ctx:
            loss = loss_fn(net)
        loss.backward()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        opt.step(closure)
    sync()
    calls.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    last = None
    for _ in range(a.steps):
        last = opt.step(closure)
    e1.record()
    sync()
    wall = time.perf_counter() - t0
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / a.steps
    if rank == 0:
        print(json.dumps({
            "metric": f"{a.model}_train_{unit.replace('/', '_per_')}", "value": round(per_step * world / (ms_per_step / 1e3), 1),
            "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms_per_step, 3),
            "wall_ms_per_step": round(1e3 * wall / a.steps, 3), "higher_is_better": True, "scaling": "weak",
            "dtype": "f32", "data": "synthetic", "loss": float(last),
            "peak_memory_gib": round(torch.cuda.max_memory_allocated() / 2**30, 2),
            "packed_activations": bool(a.packed_activations),
            "config": {"workload": f"{a.model} random-init, synthetic batch {a.batch}/GPU" +
                       (f" {a.image}x{a.image}" if a.model.startswith("resnet") else f" seq {a.seq}") +
                       f", --compress {a.compress} on {a.only}", "optimizer": type(inner).__name__,
                       "batched_optimizer_side": not a.no_batched_optimizer},
            "codec_calls_per_step": {k: v // a.steps for k, v in sorted(calls.items(), key=lambda kv: str(kv[0]))},
        }))
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                opt.step(closure)
            torch.cuda.synchronize()
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        busy = sum(e.device_time for e in ev) / 3e3
        ours = sum(e.device_time for e in ev if "smaq" in e.name or "floatq" in e.name) / 3e3
        print(f"[profile] GPU busy {busy:.3f} ms/step of {ms_per_step:.3f}; codec kernels {ours:.3f} ms/step; "
              f"{len(ev) // 3} GPU ops/step", file=sys.stderr)
        agg = Counter()
        cnt = Counter()
        for e in ev:
            agg[e.name[:70]] += e.device_time / 3e3
            cnt[e.name[:70]] += 1
        for k, v in agg.most_common(14):
            print(f"[profile] {v:8.3f} ms  x{cnt[k] // 3:<4d} {k}", file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
