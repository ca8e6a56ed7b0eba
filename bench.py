#!/usr/bin/env python
"""Benchmark of the SmaQ compress->decompress hot path on B200 (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--log2n 30]
    torchrun --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU; weak scaling)

A *step* is one pass of the hot path over one synthetic tensor of 2^log2n fp32 elements
(BASELINE.md §4 recipe: N(0,1), seed 1234, 1 % of the elements x10), with the reference's default
flags (6/8 bits, thresholds 1.0/2.5, stochastic rounding, full-tensor statistics):

    statistics kernel  ->  quantise + pack (smaq_encode)  ->  unpack + de-normalise (smaq_decode)

`value` is whole-job algorithmic GB/s with the input resident in HBM: (12 N + 2 C) bytes per step
and rank (4N statistics read, 4N + C encode, C + 4N decode; C = packed payload bytes) divided by
the device time of the K timed steps (CUDA events, max over ranks).  `e2e` is the same quantity
through the plugin API with HOST buffers: pinned host -> device copy of the input, SmartFP.encode,
SmartFP.decode, device -> pinned host copy of the result, all inside the timed region.
`roofline` is the encode kernel (the dominant one) against the measured HBM copy bandwidth in
MEASURED_PEAKS.json.  `cpu_baseline` / `--impl reference` time the CPU port of the reference's own
torch implementation of the path (oracle/smaq.py: the same operator sequence as
smart_compress/compress/smart.py:110-190, which is how the reference runs on CPU) on a bounded
sample, credited with the same algorithmic bytes per element so ratios are element-rate ratios.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "smart-quantization_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CPU_SAMPLE_LOG2N = 26  # configs[0]: 64 Mi elements, the reference's own CPU-runnable case


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=30, help="elements per tensor = 2^log2n (default 1 Gi)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the size / codec sweep (N=1 only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the ResNet-34 DDP training block")
    ap.add_argument("--train-steps", type=int, default=50, help="timed training steps (BASELINE.md §4: >= 50)")
    ap.add_argument("--train-warmup", type=int, default=10, help="untimed training steps before them (>= 10)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def max_over_ranks(elapsed_ms: float, world: int, device) -> float:
    """Device time of the slowest rank (every rank gets the same number).  The path shards with no
    exchange: this all-reduce and the barriers are the only collectives of the bench."""
    if world <= 1:
        return float(elapsed_ms)
    import torch.distributed as dist

    t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_gbs(world: int, bytes_per_rank_step: float, ms_per_step: float) -> float:
    """Weak scaling: every rank processes its own tensor; the job's throughput is the sum over ranks
    of the bytes one step moves, over the slowest rank's time."""
    return world * bytes_per_rank_step / (ms_per_step * 1e-3) / 1e9


def traffic_from_profile(kernel: str, log2n: int):
    """DRAM bytes of one launch of `kernel` from the committed ncu capture (tools/make_profiles.py)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)[kernel]
    except (OSError, KeyError, ValueError):
        return None
    return rec["dram_bytes"] if rec.get("log2n") == log2n else None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def bytes_per_element(outlier_fraction: float) -> dict:
    """Algorithmic bytes per element, BASELINE.md §4: C = (6 n_main + 8 n_out)/8 bytes."""
    c = (6.0 * (1.0 - outlier_fraction) + 8.0 * outlier_fraction) / 8.0
    return {"stats": 4.0, "encode": 4.0 + c, "decode": c + 4.0, "step": 12.0 + 2.0 * c, "roundtrip": 8.0,
            "fp8": 8.0, "s2fp8": 12.0, "packed": c}


# ------------------------------------------------------------------------------------------------------
# clocks during the timed region
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = get_reasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.002)  # the timed region of the kernel bench lasts ~30 ms

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
            # the first NVML query of a process can take longer than the kernel bench's whole timed region (one
            # sample in a 25 ms region was observed): let the thread complete one query, then drop what it read
            # before the region started
            t = time.perf_counter()
            while not self.samples and time.perf_counter() - t < 1.0:
                time.sleep(0.001)
            self.samples.clear()
            self.reasons.clear()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (its NUMA node), so that the pinned host
    buffers of the end-to-end leg are allocated next to the GPU's PCIe root.  At N = 8 the ranks otherwise share
    whichever node the launcher started them on.  Returns a short description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return "unchanged (NVML reported no local CPUs inside this process's affinity)"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs local to GPU {index} (first {min(cpus)}, last {max(cpus)})"
    except Exception as e:  # no NVML / containers that forbid sched_setaffinity: leave the placement to the OS
        return f"unchanged ({type(e).__name__})"


def make_input(n: int, device, seed=1234):
    """BASELINE.md §4 recipe, generated on the device the tensor lives on."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, generator=g, device=device, dtype=torch.float32)
    if n <= (1 << 27):
        idx = torch.randperm(n, generator=g, device=device)[: n // 100]
    else:  # randperm(2^30) would need 8 GiB + a sort; distinctness of 1 % of the indices is immaterial here
        idx = torch.randint(0, n, (n // 100,), generator=g, device=device)
    x[idx] *= 10
    return x


def make_plugin():
    from argparse import ArgumentParser

    from smart_compress.compress.smart import SmartFP

    args = SmartFP.add_argparse_args(ArgumentParser()).parse_args([])  # the reference's defaults
    args.precision = 32
    return SmartFP(args)


class Pipeline:
    """stats -> encode -> decode through the C ABI on preallocated device buffers."""

    def __init__(self, n, device):
        import ctypes as C

        from smart_compress import _native as N
        from smart_compress.compress.packed import packed_layout

        self.C, self.N, self.lib = C, N, N.load()
        self.n, self.device = n, device
        self.fp = make_plugin()
        self.lay = packed_layout(n, 6, 8)
        self.ms = torch.empty(2, dtype=torch.float32, device=device)
        self.stats_ws_bytes = self.lib.smaq_stats_workspace_bytes(n)
        self.stats_ws = torch.empty(self.stats_ws_bytes, dtype=torch.uint8, device=device)
        self.packed = torch.empty(self.lay.total_capacity_bytes, dtype=torch.uint8, device=device)
        self.enc_ws = torch.zeros(self.lay.workspace_bytes, dtype=torch.uint8, device=device)  # zeroed once; every call leaves it zero
        self.y = torch.empty(n, dtype=torch.float32, device=device)
        # statistics; encode (one pass: quantise + pack); decode
        self.launches_per_step = 3
        self.step_index = 0

    def stats(self, x):
        N = self.N
        N.check(self.lib.smaq_stats_full(x.data_ptr(), self.n, 1, self.ms.data_ptr(), self.stats_ws.data_ptr(),
                                         self.stats_ws_bytes, N.stream_ptr(self.device)), "stats")

    def encode(self, x, count_saturated=False):
        N = self.N
        params = self.fp._params(all_positive=False)
        params.count_saturated = int(count_saturated)  # plugin default: only under --measure_compression_ratio
        N.check(self.lib.smaq_encode(x.data_ptr(), self.n, self.ms.data_ptr(), None, self.C.byref(params),
                                     self.packed.data_ptr(), self.packed.numel(), self.enc_ws.data_ptr(),
                                     self.enc_ws.numel(), N.stream_ptr(self.device)), "encode")

    def decode(self):
        N = self.N
        N.check(self.lib.smaq_decode(self.packed.data_ptr(), self.packed.numel(), self.n, 6, 8, 0, self.y.data_ptr(),
                                     N.stream_ptr(self.device)), "decode")

    def header(self):
        raw = bytes(self.packed[: self.C.sizeof(self.N.PackedHeader)].cpu().numpy())
        return self.N.PackedHeader.from_buffer_copy(raw)


def time_kernel(fn, iters=5, warmup=2, reps=1):
    """Median device time (ms) of fn() with CUDA events on the current stream.  reps > 1: that many calls are
    queued between the two events and the time divided — for launches so short that a pair of events around
    ONE of them would mostly time the events."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b) / reps)
    return statistics.median(times), min(times)


def run_b200(args):
    import torch.distributed as dist

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = 1 << args.log2n
    peak, peak_src = load_peaks()
    x = make_input(n, device)
    pipe = Pipeline(n, device)

    # ---- device-resident steps ------------------------------------------------------------------------
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(record=None):
        if record is not None:
            e = [ev() for _ in range(4)]
            e[0].record()
            pipe.stats(x)
            e[1].record()
            pipe.encode(x)
            e[2].record()
            pipe.decode()
            e[3].record()
            record.append(e)
        else:
            pipe.stats(x)
            pipe.encode(x)
            pipe.decode()

    for _ in range(args.warmup):
        step()
    barrier()
    per_kernel = []
    with ClockSampler(local) as clocks:
        t0, t1 = ev(), ev()
        t0.record()
        for _ in range(args.steps):
            step(per_kernel)
        t1.record()
        barrier()
    elapsed_ms = max_over_ranks(t0.elapsed_time(t1), world, device)
    ms_per_step = elapsed_ms / args.steps

    pipe.encode(x, count_saturated=True)  # untimed: the timed steps run with the plugin's default (no count)
    hdr = pipe.header()
    assert hdr.status == 0 and hdr.n == n, "encode reported a failure"
    f_out = hdr.n_outlier / n
    bpe = bytes_per_element(f_out)
    step_bytes = bpe["step"] * n
    value = whole_job_gbs(world, step_bytes, ms_per_step)

    k_ms = {name: statistics.mean(e[i].elapsed_time(e[i + 1]) for e in per_kernel)
            for i, name in enumerate(("stats", "encode", "decode"))}
    k_gbs = {name: bpe[name] * n / (k_ms[name] * 1e-3) / 1e9 for name in k_ms}
    enc_dec_gbs = (bpe["encode"] + bpe["decode"]) * n / ((k_ms["encode"] + k_ms["decode"]) * 1e-3) / 1e9
    dominant = max(("encode", "decode"), key=lambda k: k_ms[k])

    result = {
        "metric": "smaq_encode_decode_gbs",
        "value": round(value, 1),
        "unit": "GB/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": f"SmaQ 6/8-bit stats+encode+decode of one 2^{args.log2n}-element fp32 tensor per GPU "
                        "(BASELINE configs[1], largest size of the sweep)",
            "elements_per_gpu": n,
            "input": "N(0,1) seed 1234, 1% of elements x10",
            "flags": "reference defaults: --num_bits_main 6 --num_bits_outlier 8, thresholds 1.0/2.5, "
                     "stochastic rounding (in-kernel Philox), full-tensor statistics",
            "outlier_fraction": round(f_out, 5),
            "algorithmic_bytes_per_element": {k: round(v, 4) for k, v in bpe.items()},
            "l2": f"working set {(4 * n * 2 + pipe.lay.total_capacity_bytes) / 2**30:.1f} GiB per step >> 126 MB L2 "
                  "(no flush needed)" if args.log2n >= 27 else "working set may fit in L2",
            "parallelism": f"{world} independent replicas, no data-path collective (SURVEY.md §8e)",
        },
        "gpu_launches": args.steps * pipe.launches_per_step,
        "clocks": clocks.summary(),
        "kernels": {
            name: {"ms": round(k_ms[name], 4), "gbs": round(k_gbs[name], 1), "frac_of_peak": round(k_gbs[name] / peak, 4)}
            for name in k_ms
        },
        "encode_decode_gbs": round(enc_dec_gbs, 1),
        "encode_decode_frac_of_measured_peak": round(enc_dec_gbs / peak, 4),
        "encode_decode_frac_of_nominal_8000": round(enc_dec_gbs / 8000.0, 4),
        "frac_of_hbm_peak": round(value / world / peak, 4),
        "frac_of_nominal_8000": round(value / world / 8000.0, 4),
        "roofline": {
            "bound": "hbm",
            "kernel": f"smaq::{dominant}_kernel<5, 2, ...>",
            "achieved": round(k_gbs[dominant], 1),
            "peak": peak,
            "peak_source": peak_src,
            "unit": "GB/s",
            "frac": round(k_gbs[dominant] / peak, 4),
            "frac_of_nominal_8000": round(k_gbs[dominant] / 8000.0, 4),
            "traffic": traffic_from_profile(dominant, args.log2n),
            "traffic_source": "profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one "
                              "`ncu --set full` launch at the same size (null when the size differs)",
            "algorithmic_bytes_per_launch": int(bpe[dominant] * n),
        },
        "packed": {"payload_bits_per_element": round(8 * bpe["packed"], 4),
                   "compression_ratio": round(32.0 / (8 * bpe["packed"]), 4),
                   "n_saturated": int(hdr.n_saturated)},
    }

    # ---- end to end through the plugin API with host buffers ---------------------------------------------
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 20))
        # the pinned pages are first-touched on the GPU's own NUMA node (a box whose launcher started the process
        # on the other socket moved them over the inter-socket link: 115 instead of 94-98 ms per step); at N = 1 the
        # affinity is restored after this leg — the CPU baseline later in this process wants every core
        affinity_before = os.sched_getaffinity(0)
        numa = bind_to_gpu_cpus(local)
        hx = torch.empty(n, dtype=torch.float32, pin_memory=True)
        hy = torch.empty(n, dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        fp = pipe.fp
        del x
        torch.cuda.empty_cache()
        # A pipelined data path, as a loader would run it: step i's host->device copy (stream s_in), its
        # statistics + encode + decode (s_comp) and the device->host read of its result (s_out) overlap with the
        # neighbouring steps' — PCIe is full duplex.  Every step still moves its own 4N bytes in and 4N bytes out
        # inside the timed region; device buffers are double-buffered and guarded by events.
        s_in, s_comp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        dxs = [torch.empty(n, dtype=torch.float32, device=device) for _ in range(2)]
        dys = [pipe.y, torch.empty(n, dtype=torch.float32, device=device)]
        done_in, done_comp, done_out = {}, {}, {}

        def e2e_step(i):
            b = i & 1
            with torch.cuda.stream(s_in):
                if i - 2 in done_comp:
                    s_in.wait_event(done_comp[i - 2])    # dxs[b] was the input of step i-2
                dxs[b].copy_(hx, non_blocking=True)       # host -> device, pinned
                done_in[i] = s_in.record_event()
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(done_in[i])
                if i - 2 in done_out:
                    s_comp.wait_event(done_out[i - 2])   # dys[b] was read back by step i-2
                packed = fp.encode(dxs[b])                # statistics + quantise + pack
                fp.decode(packed, out=dys[b])             # unpack + de-normalise
                done_comp[i] = s_comp.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(done_comp[i])
                hy.copy_(dys[b], non_blocking=True)       # device -> host, pinned
                done_out[i] = s_out.record_event()

        def run(first, count):
            for i in range(first, first + count):
                e2e_step(i)
            torch.cuda.current_stream().wait_event(done_out[first + count - 1])

        run(0, 2)
        barrier()
        t0, t1 = ev(), ev()
        t0.record()
        for st in (s_in, s_comp, s_out):
            st.wait_event(t0)
        run(2, e2e_steps)
        t1.record()
        barrier()
        e_ms = max_over_ranks(t0.elapsed_time(t1), world, device)
        e_ms /= e2e_steps
        result["e2e"] = {
            "value": round(world * step_bytes / (e_ms * 1e-3) / 1e9, 2),
            "unit": "GB/s",
            "h2d_bytes_per_step": 4 * n,
            "d2h_bytes_per_step": 4 * n,
            "ms_per_step": round(e_ms, 3),
            "steps": e2e_steps,
            "api": "SmartFP.encode / SmartFP.decode (smart_compress.compress.smart) on pinned host buffers; three "
                   "streams (copy in / compute / copy out), double-buffered: consecutive steps overlap on the full-duplex "
                   "PCIe link",
            "checksum": float(hy[:: max(1, n // 4096)].double().sum()),
            "host_numa_binding": numa,
        }
        x = dxs[0]
        del hx, hy, dys
        if world == 1:
            os.sched_setaffinity(0, affinity_before)

    # ---- size / codec sweep and the CPU baseline: rank 0 at N = 1 only ---------------------------------------
    if world == 1 and not args.no_sweep:
        result["sweep"] = sweep(args, device, peak)
    if world == 1 and not args.no_cpu:
        result["cpu_baseline"] = cpu_baseline(bpe["step"], steps=3, warmup=1)
        result["cpu_baseline"]["same_port_torch_eager_on_this_gpu"] = eager_gpu_baseline(bpe["step"], device)

    # ---- the other half of BASELINE.json's metric: ResNet-34 img/s at this many GPUs -------------------------------
    if not args.no_train:
        del pipe
        x = None
        try:   # the end-to-end leg's device buffers (defined only if it ran)
            dxs.clear()
            done_in.clear(), done_comp.clear(), done_out.clear()
        except NameError:
            pass
        torch.cuda.empty_cache()
        result["train"] = train_block(args, device, world, local)
        result["gpu_launches_note"] = "gpu_launches counts the kernel bench's timed region only (statistics + encode + decode per step)"

    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_block(args, device, world, local):
    """BASELINE.json configs[3] (the training half of `metric`): ResNet-34 with the reference's stem on synthetic
    224x224 images, batch 32 per GPU, SGD, fp32, random init; DDP over NCCL when world > 1 (its bucketed gradient
    all-reduce is the only collective, issued before the optimizer-side compression: reference optimizer.py:135-141).
    Three legs through the SAME hooks and the same wiring (smart_compress/util/train.py; reference util/train.py:
    197-213, models/base.py:137-163):
      img_per_s                      --compress smart on all five data structures, this library's kernels;
      plain_img_per_s                --no_compress (no hooks, the bare optimizer);
      reference_eager_cuda_img_per_s the reference's own SmartFP.__call__ as torch eager CUDA operators (the CPU
                                     oracle's port run on the GPU: a baseline leg, fewer steps — it is ~10x slower).
    >= 50 timed steps after 10 warm-ups; device-timed, max over ranks; clocks sampled during the
    timed region of the first leg."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from train_bench import run_training

    cfg = dict(model_name="resnet34", batch=32, image=224, device=device, world=world, local=local)
    clocks = ClockSampler(local)
    smart = run_training(compress="smart", steps=args.train_steps, warmup=args.train_warmup, clocks=clocks, **cfg)
    plain = run_training(compress="fp32", steps=args.train_steps, warmup=args.train_warmup, **cfg)
    ref_steps, ref_warm = max(3, args.train_steps // 5), max(2, args.train_warmup // 5)
    eager = run_training(compress="smart", codec="reference-eager", steps=ref_steps, warmup=ref_warm, **cfg)
    config3 = None
    if world == 1:
        # BASELINE configs[2]: ResNet-18 on CIFAR-shaped input, batch 256 — bound by the HOST when run eagerly
        # (360 codec calls + ~600 framework launches per 8 ms step), so also as one CUDA graph per step (§8 f-1)
        c3 = dict(model_name="resnet18", batch=256, image=32, device=device, world=1, local=local)
        e = run_training(compress="smart", steps=args.train_steps, warmup=args.train_warmup, **c3)
        gph = run_training(compress="smart", steps=args.train_steps, warmup=args.train_warmup, cuda_graph=True, **c3)
        pl = run_training(compress="fp32", steps=args.train_steps, warmup=args.train_warmup, **c3)
        plg = run_training(compress="fp32", steps=args.train_steps, warmup=args.train_warmup, cuda_graph=True, **c3)
        config3 = {"workload": e["workload"], "img_per_s_eager": e["value"], "img_per_s_cuda_graph": gph["value"],
                   "plain_img_per_s_eager": pl["value"], "plain_img_per_s_cuda_graph": plg["value"],
                   "codec_calls_per_step": e["codec_calls_per_step"], "steps": args.train_steps}
    # BASELINE configs[4]: BERT-base, STS-B shape (seq 128, batch 32 per GPU), AdamW incl. its moments
    c5 = dict(model_name="bert-base", batch=32, seq=128, device=device, world=world, local=local)
    b_steps, b_warm = max(10, args.train_steps // 2), max(5, args.train_warmup // 2)
    bert = run_training(compress="smart", steps=b_steps, warmup=b_warm, **c5)
    bert_plain = run_training(compress="fp32", steps=b_steps, warmup=b_warm, **c5)
    config5 = {"workload": bert["workload"], "seq_per_s": bert["value"], "plain_seq_per_s": bert_plain["value"],
               "ms_per_step": bert["ms_per_step"], "steps": b_steps, "warmup": b_warm,
               "codec_calls_per_step": bert["codec_calls_per_step"], "n_gpus": world}
    # the profiled steps run LAST, on every rank (they contain DDP's collectives): once torch.profiler has
    # initialised CUPTI every later launch of the process costs more host time — the host-bound ResNet-18 leg
    # measured 26.1 k img/s behind it against 31 k in a fresh process
    prof = run_training(compress="smart", steps=5, warmup=5, profile=True, **cfg).pop("profile", None)
    return {
        "metric": "resnet34_train_img_per_s", "unit": "img/s", "n_gpus": world,
        "img_per_s": smart["value"], "ms_per_step": smart["ms_per_step"],
        "plain_img_per_s": plain["value"], "plain_ms_per_step": plain["ms_per_step"],
        "reference_eager_cuda_img_per_s": eager["value"], "reference_eager_cuda_ms_per_step": eager["ms_per_step"],
        "speedup_vs_reference_eager_cuda": round(smart["value"] / eager["value"], 3),
        "fraction_of_plain": round(smart["value"] / plain["value"], 4),
        "steps": smart["steps"], "warmup": smart["warmup"],
        "reference_eager_cuda_steps": ref_steps, "reference_eager_cuda_warmup": ref_warm,
        "scaling": "weak", "dtype": "f32", "data": "synthetic",
        "config": {"workload": smart["workload"], "global_batch": 32 * world, "optimizer": "SGD lr 0.1 momentum 0.9",
                   "parallelism": f"dp{world}" + (" (DistributedDataParallel, NCCL all-reduce)" if world > 1 else ""),
                   "stem": "reference CIFAR stem (3x3 stride 1; SURVEY.md H10): 52.7 M feature-map elements per image"},
        "codec_calls_per_step": smart["codec_calls_per_step"],
        "peak_memory_gib": smart["peak_memory_gib"], "plain_peak_memory_gib": plain["peak_memory_gib"],
        "loss": smart["loss"], "clocks": clocks.summary(), "config3_resnet18_cifar": config3, "config5_bert_base": config5,
        "profile": None if prof is None else {k: prof[k] for k in ("gpu_busy_ms_per_step", "codec_kernels_ms_per_step",
                                                                    "nccl_kernels_ms_per_step", "gpu_ops_per_step")},
    }


def sweep(args, device, peak):
    """configs[1]: SmaQ size sweep 2^20..2^30 (fused round trip; encode; decode) and FP8 / S2FP8 at the top size."""
    import ctypes as C

    from smart_compress import _native as N
    from smart_compress.compress.packed import packed_layout
    from smart_compress.util.pytorch.quantization import make_floatq_params

    lib = N.load()
    out = {"smaq": [], "note": "median of 5 device timings after 2 warm-ups; up to 2^26 elements each timing queues 20 "
                               "launches between the events (a lone short launch would mostly time the events) — those "
                               "sizes fit in the 126 MB L2 and are partly bound by the host's launch rate"}
    fp = make_plugin()
    top = min(args.log2n, 30)
    for log2n in range(20, top + 1, 2):
        n = 1 << log2n
        x = make_input(n, device)
        y = torch.empty_like(x)
        ms = fp.statistics(x)
        params = fp._params(all_positive=False)
        lay = packed_layout(n, 6, 8)
        packed = torch.empty(lay.total_capacity_bytes, dtype=torch.uint8, device=device)
        ws = torch.zeros(lay.workspace_bytes, dtype=torch.uint8, device=device)
        sws_b = lib.smaq_stats_workspace_bytes(n)
        sws = torch.empty(sws_b, dtype=torch.uint8, device=device)
        st = N.stream_ptr(device)
        reps = 20 if log2n <= 26 else 1
        t_stats, _ = time_kernel(lambda: lib.smaq_stats_full(x.data_ptr(), n, 1, ms.data_ptr(), sws.data_ptr(), sws_b, st), reps=reps)
        t_rt, _ = time_kernel(lambda: lib.smaq_roundtrip(x.data_ptr(), y.data_ptr(), n, ms.data_ptr(), None, C.byref(params), st), reps=reps)
        # the call the training hooks make: statistics + round trip behind one entry point (dependent launches, no
        # memset node — smaq_stats_full above zeroes its ticket with one per call), 12 B/element
        cws_b = lib.smaq_compress_workspace_bytes(n)
        cws = torch.empty(cws_b, dtype=torch.uint8, device=device)
        N.check(lib.smaq_compress_workspace_init(cws.data_ptr(), cws_b, st), "smaq_compress_workspace_init")
        t_cmp, _ = time_kernel(lambda: lib.smaq_compress(x.data_ptr(), y.data_ptr(), n, None, C.byref(params),
                                                         cws.data_ptr(), cws_b, st), reps=reps)
        t_enc, _ = time_kernel(lambda: lib.smaq_encode(x.data_ptr(), n, ms.data_ptr(), None, C.byref(params),
                                                       packed.data_ptr(), packed.numel(), ws.data_ptr(), ws.numel(), st), reps=reps)
        t_dec, _ = time_kernel(lambda: lib.smaq_decode(packed.data_ptr(), packed.numel(), n, 6, 8, 0, y.data_ptr(), st), reps=reps)
        hdr = N.PackedHeader.from_buffer_copy(bytes(packed[: C.sizeof(N.PackedHeader)].cpu().numpy()))
        bpe = bytes_per_element(hdr.n_outlier / n)
        row = {"log2n": log2n}
        for name, t, key in (("stats", t_stats, "stats"), ("roundtrip", t_rt, "roundtrip"), ("encode", t_enc, "encode"),
                             ("decode", t_dec, "decode")):
            gbs = bpe[key] * n / (t * 1e-3) / 1e9
            row[name] = {"ms": round(t, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
        gbs = 12.0 * n / (t_cmp * 1e-3) / 1e9
        row["compress"] = {"ms": round(t_cmp, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
        out["smaq"].append(row)
        if log2n == top:
            r = {}
            p8 = make_floatq_params(5, 2, fp.hparams)
            t_fp8, _ = time_kernel(lambda: lib.smaq_float_quantize(x.data_ptr(), y.data_ptr(), n, None, C.byref(p8), st))
            mm = torch.empty(2, dtype=torch.float32, device=device)
            t_s2s, _ = time_kernel(lambda: lib.smaq_s2fp8_stats(x.data_ptr(), n, mm.data_ptr(), sws.data_ptr(), sws_b, st))
            t_s2a, _ = time_kernel(lambda: lib.smaq_s2fp8_apply(x.data_ptr(), y.data_ptr(), n, mm.data_ptr(), None, C.byref(p8), st))
            for name, t, b in (("fp8", t_fp8, 8.0), ("s2fp8_stats", t_s2s, 4.0), ("s2fp8_apply", t_s2a, 8.0),
                               ("s2fp8", t_s2s + t_s2a, 12.0)):
                gbs = b * n / (t * 1e-3) / 1e9
                r[name] = {"ms": round(t, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
            r["log2n"] = log2n
            out["float_emulation"] = r
            # also pure N(0,1) (31.7 % outliers) and --use_sample_stats (k = 16)
            g = torch.Generator(device=device).manual_seed(4321)
            xn = torch.randn(n, generator=g, device=device, dtype=torch.float32)
            N.check(lib.smaq_stats_full(xn.data_ptr(), n, 1, ms.data_ptr(), sws.data_ptr(), sws_b, st), "stats")
            t_rt_n, _ = time_kernel(lambda: lib.smaq_roundtrip(xn.data_ptr(), y.data_ptr(), n, ms.data_ptr(), None, C.byref(params), st))
            t_enc_n, _ = time_kernel(lambda: lib.smaq_encode(xn.data_ptr(), n, ms.data_ptr(), None, C.byref(params),
                                                             packed.data_ptr(), packed.numel(), ws.data_ptr(), ws.numel(), st))
            t_dec_n, _ = time_kernel(lambda: lib.smaq_decode(packed.data_ptr(), packed.numel(), n, 6, 8, 0, y.data_ptr(), st))
            hdr_n = N.PackedHeader.from_buffer_copy(bytes(packed[: C.sizeof(N.PackedHeader)].cpu().numpy()))
            bn = bytes_per_element(hdr_n.n_outlier / n)
            pn = {"log2n": log2n, "outlier_fraction": round(hdr_n.n_outlier / n, 5)}
            for name, t, key in (("roundtrip", t_rt_n, "roundtrip"), ("encode", t_enc_n, "encode"), ("decode", t_dec_n, "decode")):
                gbs = bn[key] * n / (t * 1e-3) / 1e9
                pn[name] = {"ms": round(t, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3), "frac_of_nominal_8000": round(gbs / 8000.0, 3)}
            gbs = (bn["encode"] + bn["decode"]) * n / ((t_enc_n + t_dec_n) * 1e-3) / 1e9
            pn["encode_decode"] = {"gbs": round(gbs, 1), "frac": round(gbs / peak, 3), "frac_of_nominal_8000": round(gbs / 8000.0, 3)}
            out["pure_normal_input"] = pn
            del xn
            t_ss, _ = time_kernel(lambda: lib.smaq_stats_sampled_draw(x.data_ptr(), n, 16, 0, 1234, 7, ms.data_ptr(), st), reps=20)
            N.check(lib.smaq_stats_sampled_draw(x.data_ptr(), n, 16, 0, 1234, 7, ms.data_ptr(), st), "sampled")
            t_rt_s, _ = time_kernel(lambda: lib.smaq_roundtrip(x.data_ptr(), y.data_ptr(), n, ms.data_ptr(), None, C.byref(params), st))
            gbs = 8.0 * n / ((t_ss + t_rt_s) * 1e-3) / 1e9
            out["use_sample_stats"] = {"log2n": log2n, "num_samples": 16, "stats_ms": round(t_ss, 4), "roundtrip_ms": round(t_rt_s, 4),
                                       "gbs": round(gbs, 1), "frac": round(gbs / peak, 3), "frac_of_nominal_8000": round(gbs / 8000.0, 3),
                                       "note": "k indices drawn on the device (Philox + Floyd) instead of randperm(N): 8 B/element algorithmic"}
        del x, y, packed, ws
    return out


# ------------------------------------------------------------------------------------------------------
def cpu_step_fn(log2n):
    """One 'step' of the reference's CPU implementation of the path on a 2^log2n sample: the oracle port
    (oracle/smaq.py restates smart.py:110-190 with the same torch CPU operators, incl. rand_like)."""
    from oracle.smaq import SmaqConfig, smaq_roundtrip

    n = 1 << log2n
    x = make_input(n, torch.device("cpu"))
    cfg = SmaqConfig()

    def fn():
        probs = torch.rand_like(x)  # smart.py:94 — part of the reference's per-call cost
        return smaq_roundtrip(x, cfg, probs=probs).y

    return fn, n


def cpu_baseline(step_bytes_per_element, steps, warmup):
    threads = torch.get_num_threads()
    fn, n = cpu_step_fn(CPU_SAMPLE_LOG2N)
    for _ in range(warmup):
        fn()
    times = []
    for _ in range(steps):
        t = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t)
    best = min(times)
    return {
        "value": round(step_bytes_per_element * n / best / 1e9, 4),
        "unit": "GB/s",
        "cores": threads,
        "kind": "port",
        "sample": f"2^{CPU_SAMPLE_LOG2N} elements (BASELINE configs[0]), same input recipe, best of {steps} after "
                  f"{warmup} warm-up; torch CPU ops in the reference's order incl. rand_like; credited "
                  f"{step_bytes_per_element:.3f} algorithmic B/element like the GPU step",
        "ms_per_step": round(best * 1e3, 1),
        "host_cpus": os.cpu_count(),
    }


def eager_gpu_baseline(step_bytes_per_element, device, log2n=26, iters=5):
    """Context, part of the baseline leg: the SAME port of smart.py:110-190 (oracle/smaq.py, ~26 eager torch
    operators + rand_like + the `if std == 0` host sync) evaluated by torch's own CUDA kernels on this GPU —
    what running the reference unchanged on a B200 would cost.  A reported baseline, never the product path."""
    from oracle.smaq import SmaqConfig, smaq_roundtrip

    n = 1 << log2n
    x = make_input(n, device)
    cfg = SmaqConfig()

    def fn():
        return smaq_roundtrip(x, cfg, probs=torch.rand_like(x)).y

    ms, best = time_kernel(fn, iters=iters, warmup=2)
    return {"ms_per_call": round(ms, 3), "value": round(step_bytes_per_element * n / (ms * 1e-3) / 1e9, 1), "unit": "GB/s",
            "sample": f"2^{log2n} elements, median of {iters}; torch eager CUDA operators in the reference's order, credited "
                      f"{step_bytes_per_element:.3f} algorithmic B/element like the GPU step"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or torch.get_num_threads()
    torch.set_num_threads(threads)
    fn, n = cpu_step_fn(CPU_SAMPLE_LOG2N)
    from oracle.smaq import SmaqConfig, smaq_roundtrip

    res = smaq_roundtrip(make_input(n, torch.device("cpu")), SmaqConfig(stochastic_rounding=False))
    f_out = float((res.hi | res.lo).float().mean())
    bpe = bytes_per_element(f_out)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    value = bpe["step"] * n / dt / 1e9
    line = {
        "impl": "reference",
        "metric": "smaq_encode_decode_gbs",
        "value": round(value, 4),
        "unit": "GB/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 2),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": "SmaQ 6/8-bit compress->decompress, reference defaults, CPU port of smart.py:110-190 "
                        f"(torch CPU ops) on a bounded sample of 2^{CPU_SAMPLE_LOG2N} elements per step",
            "input": "N(0,1) seed 1234, 1% of elements x10",
            "algorithmic_bytes_per_element": round(bpe["step"], 4),
        },
        "cpu_baseline": {"value": round(value, 4), "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"2^{CPU_SAMPLE_LOG2N} elements per step, {args.steps} steps after {args.warmup} warm-up"},
        "e2e": {"value": round(value, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
