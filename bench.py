#!/usr/bin/env python
"""Benchmark of the EAMM per-frame generation hot path on B200 (contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU)
    python bench.py --impl reference ...                     # the reference algorithm on the host CPUs

A "step" is one pass of DenseMotionNetwork + OcclusionAwareGenerator over one batch of synthetic
256x256 frames (BASELINE.json configs[1]: batch 32, 10 keypoints; fp32-equivalent arithmetic by
default).  One JSON line is printed by rank 0.  Work per GPU is fixed (weak scaling): with N ranks
the job processes N*batch frames per step, frames block-partitioned, no data-path collective.

Extra keys of the same line (each a separately timed leg after the headline region):
  cudnn_baseline   the reference's own GPU path (demo.py:98-109, :279): the same algorithm as eager PyTorch-CUDA ops
                   (cuDNN / ATen kernels, fp32, allow_tf32 off and on) on this GPU -- the stronger second bar next to
                   the CPU arm (N = 1 only)
  configs          BASELINE.json configs[2] (batch 256, one-pass reduced-precision convs / fp32 warp; N = 1) and
                   configs[3] (1024 frames block-partitioned over the ranks, i.e. 128 per GPU at N = 8, plus the
                   optional NCCL gather of the uint8 frames onto rank 0; N > 1)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from eamm_b200 import get_config, synth, sharding           # noqa: E402

ALG_GFLOP_PER_FRAME = 107.286       # 30 convs, 2*MAC, per (source, kp) frame
# ncu --set full DRAM traffic (read + write bytes) of ONE bottleneck-conv launch at B=32: (bytes, source file)
TRAFFIC = {"fp32_bf16x3": (138.35e6 + 95.15e6, "profiles/r1b_ncu_summary.md"),
           # mixed fp16 + 2 x e4m3 operands: mean of conv1 (136.7 + 91.6 MB) and conv2 (278.0 + 227.6 MB: residual read, two outputs)
           "mix": (0.5 * (136.70e6 + 91.59e6 + 278.02e6 + 227.63e6), "profiles/r2_ncu_final.md (res_conv1 + res_conv2, mean)")}
METRIC = "256x256 frames/sec (DenseMotionNetwork + OcclusionAwareGenerator forward, 10 kp)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU set local to GPU `index` (NVML affinity) before any pinned host
    buffer is allocated: first-touch then places the staging buffers on the GPU's NUMA node, which
    keeps the H2D/D2H copies of the e2e path at PCIe speed instead of crossing the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    try:        # NVML affinity unavailable (containers): the PCI device's local_cpulist from sysfs
        pr = torch.cuda.get_device_properties(index)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = []
        with open(path) as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def cpu_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return model


def oracle_frames_per_sec(batch, repeats, warmup, threads=None):
    """Reference algorithm (CPU oracle = the reference's own torch CPU ops) timed on the host cores."""
    from oracle import eamm_oracle as oracle       # the one place bench.py executes oracle/ (as baseline)
    if threads:
        torch.set_num_threads(threads)
    cfg = get_config("full")
    sd = synth.make_state_dict(cfg, seed=0)
    src, kpd, kps = synth.make_inputs(batch, cfg, size=256, seed=1)
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        oracle.generator_forward(sd, cfg, src, kpd, kps)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch / statistics.median(times), times


def _time_steps(fn, steps, warmup, barrier=None):
    """ms per step of fn(), CUDA events on the current stream, after warm-up, synchronised on both sides."""
    for _ in range(warmup):
        fn()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def cudnn_baseline(cfg, dev, B, d_src, d_kpd, d_kps, our_value):
    """The reference's own GPU path (demo.py:98-109 `.cuda().eval()`, :279) on this GPU: the same algorithm as eager
    PyTorch-CUDA ops -- cuDNN / ATen kernels, one per op, fp32 -- with torch.backends.cudnn.allow_tf32 off and on.
    (The oracle functions are device-agnostic torch.nn.functional calls; here their tensors live on cuda:0.  This leg
    is a reported baseline: none of our kernels run in it.)"""
    from oracle import eamm_oracle as oracle          # baseline leg only
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(cfg, seed=0).items()}
    src = d_src.contiguous()
    out = {"unit": "frames/s", "batch": B, "what": "eager PyTorch-CUDA (cuDNN/ATen) run of the reference algorithm, fp32 storage"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                ms = _time_steps(lambda: oracle.generator_forward(sd, cfg, src, d_kpd, d_kps), 10, 3)
            out[name] = {"value": B / (ms * 1e-3), "ms_per_step": ms, "allow_tf32": tf32,
                         "ours_over_this": our_value / (B / (ms * 1e-3))}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    del sd
    torch.cuda.empty_cache()
    return out


def config2_leg(gen, cfg, dev, peaks, batch=256, steps=6):
    """BASELINE.json configs[2]: batch 256, 256x256, one-pass reduced-precision convs / fp32 warp, 1 GPU."""
    from eamm_b200 import engine
    res = {"workload": "BASELINE configs[2]: batch %d frames, 256x256, reduced-precision conv / fp32 warp, 1 GPU" % batch}
    src, kpd, kps = synth.make_inputs(batch, cfg, size=256, seed=2)
    d_src = src.to(dev)
    d_kpd = {k: v.to(dev) for k, v in kpd.items()}
    d_kps = {k: v.to(dev) for k, v in kps.items()}
    old = gen.precision
    try:
        for prec in ("fp16", "bf16"):
            gen.precision = prec
            ms = _time_steps(lambda: gen(d_src, kp_driving=d_kpd, kp_source=d_kps), steps, 3)
            engine.PROFILE = []
            if hasattr(torch.cuda, "_sleep"):
                torch.cuda._sleep(20_000_000)
            gen(d_src, kp_driving=d_kpd, kp_source=d_kps)
            torch.cuda.synchronize()
            prof, engine.PROFILE = engine.PROFILE, None
            dom = [(fl, a.elapsed_time(b)) for name, fl, nb, a, b in prof if name.startswith("conv:res")]
            wo = [(nb, a.elapsed_time(b)) for name, fl, nb, a, b in prof if name == "warp_occlude"]
            tf = sum(f for f, _ in dom) / (sum(t for _, t in dom) * 1e-3) / 1e12
            res[prec] = {"value": batch / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "steps": steps,
                         "bottleneck_conv_tflops": tf, "bottleneck_conv_frac_of_peak": tf / peaks["bf16_tflops_sustained"],
                         "warp_occlude_gbs": wo[0][0] / (wo[0][1] * 1e-3) / 1e9 if wo else None}
            gen._eng.ws.clear(); gen._eng.dm.ws.clear()
            torch.cuda.empty_cache()
    finally:
        gen.precision = old
    return res


def config3_leg(gen, cfg, dev, rank, world, total=1024, steps=5):
    """BASELINE.json configs[3]: 1024 frames block-partitioned over the ranks (128 per GPU at N = 8), frames/s of the
    whole job, plus the optional tail of SURVEY 8(e): NCCL gather of the uint8 frames onto rank 0."""
    import torch.distributed as dist
    start, stop = sharding.partition(total, world, rank)
    n = stop - start
    src_all, kpd_all, kps_all = synth.make_inputs(total, cfg, size=256, seed=3)
    sl = slice(start, stop)
    d_src = src_all[sl].contiguous().to(dev)
    d_kpd = {k: v[sl].contiguous().to(dev) for k, v in kpd_all.items()}
    d_kps = {k: v[sl].contiguous().to(dev) for k, v in kps_all.items()}
    del src_all
    gen.emit_u8 = True
    try:
        last = {}

        def step():
            last["out"] = gen(d_src, kp_driving=d_kpd, kp_source=d_kps)

        ms = sharding.reduce_max(_time_steps(step, steps, 3, barrier=dist.barrier), device=dev)
        frames_u8 = last["out"]["prediction_u8"]

        def step_gather():
            step()
            last["gathered"] = sharding.gather_frames(last["out"]["prediction_u8"], total)

        ms_g = sharding.reduce_max(_time_steps(step_gather, steps, 2, barrier=dist.barrier), device=dev)
        ok = True
        if rank == 0:
            g = last["gathered"]
            ok = tuple(g.shape) == (total, 256, 256, 3) and bool(torch.equal(g[start:stop], frames_u8))
    finally:
        gen.emit_u8 = False
    gen._eng.ws.clear(); gen._eng.dm.ws.clear()
    torch.cuda.empty_cache()
    return {"workload": "BASELINE configs[3]: %d frames sharded over %d GPUs (%d per GPU), precision %s" % (total, world, n, gen.precision),
            "value": total / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms, "steps": steps, "frames_per_gpu": n,
            "with_gather_u8": {"value": total / (ms_g * 1e-3), "ms_per_step": ms_g, "gather_ms": ms_g - ms,
                               "bytes_to_root": (total - n) * 256 * 256 * 3, "root_has_all_frames": ok}}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    fps, times = oracle_frames_per_sec(args.ref_batch, args.steps, args.warmup, threads)
    total = sum(times)
    sample = "each step = one batch of %d frames (same synthetic recipe) through the oracle's torch-CPU ops" % args.ref_batch
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: batch %d frames, 256x256, 10 kp, DenseMotion+Generator, distinct source "
                               "per frame (reference arm: the oracle's torch-CPU ops)" % args.ref_batch, "cpu": cpu_info()},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("EAMM_B200_PRECISION", "fp32"),
                    choices=["fp32", "fp32_bf16x3", "fp16", "bf16", "fp32_simt"])
    ap.add_argument("--shared-source", action="store_true", help="one source image for the whole batch")
    ap.add_argument("--ref-batch", type=int, default=0, help="frames per step of the CPU reference arm (0 = --batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cudnn_baseline and configs[2]/[3] legs")
    ap.add_argument("--all-kernels", action="store_true", help="kernels_ms_per_step lists every kernel, not the top 12")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.ref_batch <= 0:
        args.ref_batch = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch.distributed as dist
    from eamm_b200 import engine, _lib
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = get_config("full")
    gen = OcclusionAwareGenerator(**cfg).eval()
    gen.load_state_dict(synth.make_state_dict(cfg, seed=0))
    gen = gen.to(dev)
    gen.precision = args.precision
    gen.strict_errors = False            # singular-Jacobian flag is checked once after the timed loop
    B = args.batch
    total = B * world
    start, stop = sharding.partition(total, world, rank)
    # every rank draws the same global batch and keeps its block (identical to a scatter of the inputs)
    src_all, kpd_all, kps_all = synth.make_inputs(total, cfg, size=256, seed=1, shared_source=args.shared_source)
    sl = slice(start, stop)
    h_src = src_all[sl].contiguous().pin_memory()
    h_kpd = {k: v[sl].contiguous().pin_memory() for k, v in kpd_all.items()}
    h_kps = {k: v[sl].contiguous().pin_memory() for k, v in kps_all.items()}
    d_src = h_src.to(dev)
    if args.shared_source:
        d_src = d_src[:1].expand(B, -1, -1, -1)
    d_kpd = {k: v.to(dev) for k, v in h_kpd.items()}
    d_kps = {k: v.to(dev) for k, v in h_kps.items()}

    def step_device():
        return gen(d_src, kp_driving=d_kpd, kp_source=d_kps)

    from eamm_b200.pipeline import FramePipeline
    h_outs = [torch.empty(B, 3, 256, 256, dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = FramePipeline(gen, depth=2)
    e2e_i = [0]

    def step_e2e():
        # public API: pinned host inputs in, frames back on the host (every step uploads its inputs and
        # downloads its result; copies overlap the neighbouring steps' compute on separate streams)
        pipe.submit(h_src, h_kpd, h_kps, h_outs[e2e_i[0] % 2])
        e2e_i[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()                                     # e.g. wait for the last frames to reach the host
        e1.record()
        barrier()
        return sharding.reduce_max(e0.elapsed_time(e1), device=dev)     # ms, max over ranks

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _lib.LAUNCHES
    ms = timed(step_device, args.steps)
    launches = _lib.LAUNCHES - l0
    clocks = sampler.stop() if sampler else None
    from eamm_b200.modules.dense_motion import check_status
    check_status(gen._eng.dm.status, clear=True)

    for _ in range(3):
        step_e2e()
    pipe.drain()
    # two passes of K steps each; the MEAN is the headline and both are listed (the host<->device copies of a pass
    # occasionally run at a fraction of PCIe speed for the whole pass, seen on ~1 run in 5 on the pool's boxes)
    e2e_passes = [timed(step_e2e, args.steps, finish=pipe.drain) for _ in range(2)]
    ms_e2e = sum(e2e_passes) / len(e2e_passes)
    pipe.close()

    # per-kernel roofline pass: one extra step with every launch bracketed by CUDA events
    # (the GPU is parked on a ~3 ms spin first, so that the host has queued the whole step before the first kernel
    #  starts: otherwise the interval around a 20-100 us kernel also holds the host's launch latency)
    engine.PROFILE = []
    if hasattr(torch.cuda, "_sleep"):
        torch.cuda._sleep(6_000_000)
    step_device()
    torch.cuda.synchronize()
    prof, engine.PROFILE = engine.PROFILE, None
    per = {}
    for name, fl, nb, a, b in prof:
        d = per.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += fl; d[2] += nb; d[3] += 1
    peaks = load_peaks()
    conv_ms = sum(v[0] for k, v in per.items() if k.startswith("conv:"))
    conv_fl = sum(v[1] for k, v in per.items() if k.startswith("conv:"))
    step_ms_prof = sum(v[0] for v in per.values())
    # dominant kernel: the 3x3 256->256 bottleneck convolutions (12 launches, 54% of the FLOPs)
    dom = [(k, v) for k, v in per.items() if k.startswith("conv:res")]
    dom_ms = sum(v[0] for _, v in dom); dom_fl = sum(v[1] for _, v in dom); dom_n = sum(v[3] for _, v in dom)
    achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    mixed = bool(getattr(gen._eng, "mixed", False))
    # bf16-pass equivalents executed per algorithmic FLOP by the bottleneck convs: 3 bf16 passes (hi/lo planes), or one
    # fp16 pass + two fp8 passes at twice the rate (mixed fp16 + 2 x e4m3 operands), or a single pass
    passes = (2 if mixed else 3) if args.precision in ("fp32", "fp32_bf16x3") else 1
    peak = peaks["bf16_tflops_sustained"]
    traffic = TRAFFIC.get("mix" if mixed else args.precision)
    roofline = {"bound": "tensor", "kernel": "conv_tc_kernel<cta_group::2> (bottleneck 3x3 256->256, %d launches/step, %s operands)"
                % (dom_n, "fp16 + 2 x e4m3" if mixed else {"fp32": "bf16 hi/lo", "fp32_bf16x3": "bf16 hi/lo"}.get(args.precision, args.precision)),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks["source"] == "measured"
                else "fallback (B200_PROFILING.md)",
                # dram__bytes_read.sum + dram__bytes_write.sum of one such launch, NOT measured by this run: the figure of
                # the ncu --set full capture named in traffic_source (at B=32), scaled with the batch
                "traffic": traffic[0] * B / 32.0 if traffic else None,
                "traffic_source": traffic[1] if traffic else None,
                "share_of_step": dom_ms / step_ms_prof if step_ms_prof else None,
                "all_convs_tflops": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
                "all_convs_share_of_step": conv_ms / step_ms_prof if step_ms_prof else None,
                # what the tensor pipe actually runs: 3 bf16 MMA passes per algorithmic FLOP in fp32 mode
                "executed_tflops": achieved * passes, "executed_frac": achieved * passes / peak,
                "pass_equivalents": passes,
                "note": "achieved/frac count the ALGORITHMIC FLOPs of the reference conv (2*MAC); the fp32-equivalent modes "
                        "execute 3 (bf16 hi/lo) or 2 (fp16 + fp8 cross terms) bf16-pass equivalents per FLOP, so frac <= 1/3 "
                        "or 1/2 there -- executed_frac is the tensor-pipe view"}
    wo = per.get("warp_occlude")
    hbm = None
    if wo:
        gbs = wo[2] / (wo[0] * 1e-3) / 1e9
        hbm = {"kernel": "warp_occlude_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
               "frac": gbs / peaks["hbm_gbs"], "bytes_per_launch": wo[2]}
    # every HBM-bound (non-conv) kernel of the step against the copy peak; the 12-45 us ones are latency-, not
    # bandwidth-limited at this batch (algorithmic bytes per launch as in SURVEY 8(d), fp32-equivalent storage)
    hbm_all = {k: {"ms": round(v[0], 4), "mbytes": round(v[2] / 1e6, 2), "gbs": round(v[2] / (v[0] * 1e-3) / 1e9, 1),
                   "frac": round(v[2] / (v[0] * 1e-3) / 1e9 / peaks["hbm_gbs"], 3)}
               for k, v in per.items() if not k.startswith("conv:") and v[0] > 0 and v[2] > 0}
    kernels = {k: {"ms": round(v[0], 4), "launches": v[3]} for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:(None if args.all_kernels else 12)]}

    frames = total * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    h2d = h_src.numel() * 4 + sum(v.numel() * 4 for v in h_kpd.values()) + sum(v.numel() * 4 for v in h_kps.values())
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"fp32": "fp32-equivalent split operands: fp16 + 2 x e4m3 (generator 3x3 / UP2 convs) / bf16 hi+lo (7x7 convs, hourglass), fp32 accumulate"
                  if mixed else "bf16x3 split (fp32-equivalent, fp32 accumulate)",
                  "fp32_bf16x3": "bf16x3 split (fp32-equivalent, fp32 accumulate)",
                  "fp16": "fp16 (fp32 accumulate, fp32 warp)", "bf16": "bf16 (fp32 accumulate, fp32 warp)",
                  "fp32_simt": "f32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: batch %d frames/GPU, 256x256, 10 kp, DenseMotion+Generator, %s"
                               % (B, "one shared source" if args.shared_source else "distinct source per frame"),
                   "precision": args.precision, "global_batch": total, "parallelism": "frames block-partitioned x%d" % world,
                   "l2": "per-step working set (activations %.1f GB + weights) exceeds the 126 MB L2; no explicit flush"
                         % (B * 0.085 if args.precision not in ("bf16", "fp16") else B * 0.043)},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": h_outs[0].numel() * 4 * world, "ms_per_step": ms_e2e / args.steps,
                "passes_ms_per_step": [round(t / args.steps, 4) for t in e2e_passes], "reported": "mean of the passes",
                "numa_local_cpus": numa_cpus},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_hbm": hbm,
        "hbm_kernels": hbm_all,
        "tensor_frac_whole_step": (ALG_GFLOP_PER_FRAME * 1e9 * frames / (ms * 1e-3) / 1e12) / peak / world,
        "kernels_ms_per_step": kernels,
    }
    if not args.no_extras:
        del pipe, h_outs
        gen._eng.ws.clear(); gen._eng.dm.ws.clear()
        torch.cuda.empty_cache()
        extras = {}

        def leg(fn, *a):
            # the extra legs are reported next to the headline, never instead of it: a failure in one of them (say, an
            # allocation that does not fit beside another tenant of the box) is recorded and the line is still printed
            try:
                return fn(*a)
            except Exception as e:                      # noqa: BLE001
                return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

        if world == 1:
            line["cudnn_baseline"] = leg(cudnn_baseline, cfg, dev, B, d_src, d_kpd, d_kps, value)
            extras["configs[2]"] = leg(config2_leg, gen, cfg, dev, peaks)
        else:
            extras["configs[3]"] = leg(config3_leg, gen, cfg, dev, rank, world)
        line["configs"] = extras
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, 64)
        try:
            fps, times = oracle_frames_per_sec(args.ref_batch, 3, 1, cores)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "3 timed batches of %d frames through the oracle (reference algorithm, "
                                              "torch CPU ops), median; cpu=%s" % (args.ref_batch, cpu_info())}
        except Exception as e:                          # noqa: BLE001  (reported baseline: never costs the headline line)
            line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300]), "kind": "port", "cores": cores}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
