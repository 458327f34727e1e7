#!/usr/bin/env python
"""Per-frame latency of the reference's real access pattern (demo.py:251-281: batch 1, one source for
the whole clip): eager calls, opt-in source cache, CUDA-graph replay, graph with a fixed source."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eamm_b200 import get_config, synth
from eamm_b200.modules.generator import OcclusionAwareGenerator
from eamm_b200.graph import GraphedGenerator

dev = torch.device("cuda:0")
cfg = get_config("full")
gen = OcclusionAwareGenerator(**cfg).eval(); gen.load_state_dict(synth.make_state_dict(cfg)); gen = gen.to(dev)
gen.strict_errors = False
src, kpd, kps = synth.make_inputs(1, cfg, size=256, seed=1)
T = 64
drv = [{k: v.to(dev) for k, v in synth.make_inputs(1, cfg, size=256, seed=10 + i)[1].items()} for i in range(T)]
s, ks = src.to(dev), {k: v.to(dev) for k, v in kps.items()}


def timed(fn):
    for i in range(8): fn(drv[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(T): fn(drv[i])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / T


for prec in ("fp32", "bf16"):
    gen.precision = prec
    gen.cache_source = False
    t_eager = timed(lambda kd: gen(s, kp_driving=kd, kp_source=ks))
    gen.cache_source = True
    t_cache = timed(lambda kd: gen(s, kp_driving=kd, kp_source=ks))
    gen.cache_source = False
    g1 = GraphedGenerator(gen, s, drv[0], ks)
    t_graph = timed(lambda kd: g1(s, kd, ks, check=False))
    g2 = GraphedGenerator(gen, s, drv[0], ks, fixed_source=True)
    t_gfix = timed(lambda kd: g2(s, kd, ks, check=False))
    print("%s B=1 ms/frame: eager %.3f | source cache %.3f | graph %.3f | graph + fixed source %.3f  (%.0f fps)"
          % (prec, t_eager, t_cache, t_graph, t_gfix, 1e3 / t_gfix))
