#!/usr/bin/env python
"""Throughput of the keypoint-detector heads on the GPU vs the CPU oracle."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eamm_b200 import synth
from eamm_b200.config import get_kp_config
from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
from oracle import eamm_oracle as oracle

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for audio in (True, False):
    cfg = get_kp_config("full", audio=audio)
    sd = synth.make_kp_state_dict(cfg, seed=3 if audio else 2)
    det = (KPDetector_a if audio else KPDetector)(**cfg).eval(); det.load_state_dict(sd); det = det.to(dev)
    x = synth.make_kp_inputs(cfg, B, 256, audio)
    xd = x.to(dev)
    for prec in ("fp32", "bf16"):
        det.precision = prec
        for _ in range(3): det(xd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): out = det(xd)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%s %s B=%d: %.3f ms/call, %.0f frames/s" % ("KPDetector_a" if audio else "KPDetector  ", prec, B, ms, B / ms * 1e3))
    torch.set_num_threads(min(64, os.cpu_count()))
    fn = oracle.kp_detector_a_forward if audio else oracle.kp_detector_forward
    fn(sd, cfg, x[:8]); t0 = time.perf_counter(); fn(sd, cfg, x[:8]); dt = time.perf_counter() - t0
    print("   CPU oracle (%d threads): %.1f frames/s" % (torch.get_num_threads(), 8 / dt))
