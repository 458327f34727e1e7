#!/usr/bin/env python
"""Per-clip time of AT_net2 on the GPU with a per-kernel breakdown, vs the CPU oracle.
usage: bench_at.py [T] [B]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eamm_b200 import synth, engine
from eamm_b200.modules.util import AT_net2
from oracle import eamm_oracle as oracle

dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sd = synth.make_at_state_dict()
m = AT_net2().eval(); m.load_state_dict(sd); m = m.to(dev)
img, mfcc, pose = synth.make_at_inputs(B, T)
a = [t.to(dev) for t in (img, mfcc, pose)]
for _ in range(3): m(*a, "cnn", 1.6)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = m(*a, "cnn", 1.6)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("AT_net2 B=%d T=%d: %.3f ms/clip, %.0f frames/s" % (B, T, ms, B * T / ms * 1e3))
engine.PROFILE = []
m(*a, "cnn", 1.6)
torch.cuda.synchronize()
rows = [(n, f, e0.elapsed_time(e1)) for n, f, _, e0, e1 in engine.PROFILE]
engine.PROFILE = None
tot = sum(r[2] for r in rows)
for n, f, t in rows:
    print("  %-18s %8.3f ms %5.1f%%  %7.2f TFLOP/s" % (n, t, 100 * t / tot, f / t / 1e9 if t > 0 else 0))
torch.set_num_threads(min(64, os.cpu_count()))
n = min(T, 16)
oracle.at_net2_forward(sd, img, mfcc[:, :n], pose[:, :n], 1.6)
t0 = time.perf_counter(); oracle.at_net2_forward(sd, img, mfcc[:, :n], pose[:, :n], 1.6); dt = time.perf_counter() - t0
print("CPU oracle (%d threads): %.1f frames/s" % (torch.get_num_threads(), B * n / dt))
