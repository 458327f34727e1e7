#!/usr/bin/env python
"""Repeat the pipelined end-to-end measurement to see its run-to-run spread (GPU box tool)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eamm_b200 import get_config, synth
from eamm_b200.modules.generator import OcclusionAwareGenerator
from eamm_b200.pipeline import FramePipeline

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda:0")
cfg = get_config("full")
gen = OcclusionAwareGenerator(**cfg).eval(); gen.load_state_dict(synth.make_state_dict(cfg)); gen = gen.to(dev); gen.precision = prec
B = 32
src, kpd, kps = synth.make_inputs(B, cfg)
h_src = src.pin_memory(); h_kpd = {k: v.pin_memory() for k, v in kpd.items()}; h_kps = {k: v.pin_memory() for k, v in kps.items()}
outs = [torch.empty(B, 3, 256, 256).pin_memory() for _ in range(2)]
print("pinned:", h_src.is_pinned(), outs[0].is_pinned())
pipe = FramePipeline(gen, depth=2)
for i in range(4): pipe.submit(h_src, h_kpd, h_kps, outs[i % 2])
pipe.drain()
for rep in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20): pipe.submit(h_src, h_kpd, h_kps, outs[i % 2])
    pipe.drain(); dt = time.perf_counter() - t0
    # copies alone
    torch.cuda.synchronize(); t1 = time.perf_counter()
    d = torch.empty_like(h_src, device=dev)
    for i in range(20): d.copy_(h_src, non_blocking=True)
    torch.cuda.synchronize(); dh = time.perf_counter() - t1
    t2 = time.perf_counter()
    for i in range(20): outs[0].copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dd = time.perf_counter() - t2
    print("rep %d: e2e %.2f ms/step (%.0f fps) | H2D %.1f GB/s  D2H %.1f GB/s" % (rep, dt / 20 * 1e3, B * 20 / dt, 20 * h_src.numel() * 4 / dh / 1e9, 20 * h_src.numel() * 4 / dd / 1e9))
