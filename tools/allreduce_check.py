#!/usr/bin/env python
"""Compressed all-reduce (smart_compress/util/pytorch/allreduce.py) on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/allreduce_check.py

For both transports (p2p: kernels read the peers' symmetric memory over NVLink; nccl: packed bytes moved by NCCL):
  * every rank ends with the SAME bits;
  * the result is the CPU ORACLE's, bit for bit: per rank the oracle's saturated round trip of its bucket (the
    kernel's statistics from the stream header, the kernel's own random numbers via oracle/rng.py), summed in rank
    order in fp32, scaled by 1/N, then the oracle's round trip of each reduced shard;
  * it is close to the exact mean (quantisation noise only);
  * timing against NCCL's fp32 all-reduce on a 32 Mi-element bucket, and the wire bytes of both.
Exit code 0 = all checks passed (rank 0 prints one JSON line)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import rng as orng
    from oracle.smaq import SmaqConfig, smaq_roundtrip
    from smart_compress import _native as N
    from smart_compress.compress.smart import SmartFP
    from smart_compress.util.pytorch.allreduce import CTA_TILE, CompressedAllReduce
    from smart_compress.util.train import parse_compression_args

    hp = parse_compression_args(["--compress", "smart"])
    cfg = SmaqConfig()
    out = {"world": world, "checks": []}
    ok = True
    for transport in ("p2p", "nccl"):
        for n in (70001, (1 << 20) + 13, 1 << 22):
            seed = 1000 + 17 * rank
            torch.manual_seed(seed)
            codec = SmartFP(hp)
            car = CompressedAllReduce(codec, n, transport=transport, min_numel=1 << 10)
            g = torch.Generator().manual_seed(n + rank)
            # gradient-like: Gaussian with a mild heavy tail (values beyond 2.5 sigma saturate in the packed format —
            # the H1 rule — which is one reason this path is opt-in)
            x = torch.randn(n, generator=g) * (1.0 + 0.1 * rank) + 0.01 * rank
            x[torch.randperm(n, generator=g)[: n // 200]] *= 2
            xd = x.to(dev)
            exact = xd.clone()
            dist.all_reduce(exact)
            exact /= world
            y = car.allreduce_mean_(xd.clone())
            torch.cuda.synchronize()
            plan = car.plans[n]
            # (a) all ranks identical
            gathered = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(gathered, y)
            same = all(torch.equal(gathered[0].view(torch.int32), t.view(torch.int32)) for t in gathered)
            # (b) close to the exact mean
            err = float((y - exact).norm() / exact.norm())
            # (c) the oracle pipeline on rank 0 (inputs, seeds and stream headers gathered)
            xs = [torch.empty_like(xd) for _ in range(world)]
            dist.all_gather(xs, xd)
            hdr_a = car.arena.buf[:128].clone()
            hdr_b = car.arena.buf[car.arena.cap_a: car.arena.cap_a + 128].clone()
            hdrs_a = [torch.empty_like(hdr_a) for _ in range(world)]
            hdrs_b = [torch.empty_like(hdr_b) for _ in range(world)]
            dist.all_gather(hdrs_a, hdr_a)
            dist.all_gather(hdrs_b, hdr_b)
            oracle_ok = None
            if rank == 0:
                def header(t):
                    return N.PackedHeader.from_buffer_copy(bytes(t.cpu().numpy())[: C.sizeof(N.PackedHeader)])

                total = torch.zeros(n)
                for r in range(world):
                    h = header(hdrs_a[r])
                    probs = torch.from_numpy(orng.probs_for(n, seed=1000 + 17 * r, offset=0))
                    dec = smaq_roundtrip(xs[r].cpu(), cfg, probs=probs, mean=torch.tensor(h.mean), std=torch.tensor(h.std_raw),
                                         rng_rule=True, saturate=True).y
                    total = total + dec          # fp32, rank order
                total = total * torch.tensor(1.0 / world, dtype=torch.float32)
                want = torch.empty(n)
                for r in range(world):
                    f, c = plan.shards[r]
                    e = plan.shard_elems[r]
                    if e == 0:
                        continue
                    h = header(hdrs_b[r])
                    sl = slice(f * CTA_TILE, f * CTA_TILE + e)
                    probs = torch.from_numpy(orng.probs_for(e, seed=1000 + 17 * r, offset=1))
                    want[sl] = smaq_roundtrip(total[sl], cfg, probs=probs, mean=torch.tensor(h.mean), std=torch.tensor(h.std_raw),
                                              rng_rule=True, saturate=True).y
                oracle_ok = bool(torch.equal(want.view(torch.int32), y.cpu().view(torch.int32)))
            rec = {"transport": transport, "n": n, "ranks_identical": same, "rel_error_vs_exact_mean": round(err, 5),
                   "bit_exact_vs_oracle": oracle_ok}
            out["checks"].append(rec)
            ok = ok and same and err < 0.15 and (oracle_ok is None or oracle_ok)
            del car, plan
    # timing on a large bucket
    n = 1 << 25
    xd = torch.randn(n, device=dev)
    torch.manual_seed(5 + rank)
    for transport in ("p2p", "nccl"):
        car = CompressedAllReduce(SmartFP(hp), n, transport=transport)
        for _ in range(3):
            car.allreduce_mean_(xd.clone())
        buf = xd.clone()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            car.allreduce_mean_(buf)
        e1.record()
        torch.cuda.synchronize()
        out[f"compressed_{transport}_ms"] = round(e0.elapsed_time(e1) / 10, 3)
        out["wire_bytes_per_rank"] = car.stats["wire_bytes"] // max(car.stats["compressed_buckets"], 1)
        out["fp32_wire_bytes_per_rank"] = car.stats["fp32_wire_bytes"] // max(car.stats["compressed_buckets"], 1)
    buf = xd.clone()
    for _ in range(3):
        dist.all_reduce(buf)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    out["nccl_fp32_allreduce_ms"] = round(e0.elapsed_time(e1) / 10, 3)
    out["bucket_elements"] = n
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
