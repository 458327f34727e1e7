#!/usr/bin/env python
"""BASELINE.json config 5 on synthetic MFCC: one 300-frame audio-driven clip, MFCC windows on the host -> uint8 frames
on the host, through AT_net2 -> KPDetector / KPDetector_a -> clip glue -> generator; timing plus PSNR against the same
chain evaluated with the CPU oracles.   usage: clip_e2e.py [T] [--no-oracle] [--out file.json]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eamm_b200 import clip, get_config, synth
from eamm_b200.config import get_kp_config
from eamm_b200.modules.generator import OcclusionAwareGenerator
from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
from eamm_b200.modules.util import AT_net2

args = [a for a in sys.argv[1:] if not a.startswith("--")]
T = int(args[0]) if args else 300
dev = torch.device("cuda:0")
cfg, kcfg, acfg = get_config("full"), get_kp_config("full"), get_kp_config("full", audio=True)
sd, ksd, asd, atsd = (synth.make_state_dict(cfg, seed=0), synth.make_kp_state_dict(kcfg, seed=2),
                      synth.make_kp_state_dict(acfg, seed=3), synth.make_at_state_dict())
gen = OcclusionAwareGenerator(**cfg).eval(); gen.load_state_dict(sd); gen = gen.to(dev)
det = KPDetector(**kcfg).eval(); det.load_state_dict(ksd); det = det.to(dev)
det_a = KPDetector_a(**acfg).eval(); det_a.load_state_dict(asd); det_a = det_a.to(dev)
at = AT_net2().eval(); at.load_state_dict(atsd); at = at.to(dev)
img, mfcc, pose = synth.make_at_inputs(1, T, seed=21)
h_in = [t.pin_memory() for t in (img, mfcc, pose)]
h_out = torch.empty(T, 256, 256, 3, dtype=torch.uint8).pin_memory()
res = {"workload": "config5: 1 clip, %d frames 256x256, synthetic MFCC" % T, "frames": T}
for prec in ("fp32", "bf16"):
    gen.precision = prec
    det.precision = det_a.precision = "fp32"

    def run():
        d = [t.to(dev, non_blocking=True) for t in h_in]
        frames = clip.animate_audio_clip(at, det, det_a, gen, d[0], d[1], d[2], 1.6)
        h_out.copy_(frames, non_blocking=True)
        return frames
    for _ in range(2): run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    res[prec] = {"ms_per_clip": dt * 1e3, "frames_per_s": T / dt}
    print("%s: %.1f ms per %d-frame clip (host MFCC -> host u8 frames), %.0f frames/s" % (prec, dt * 1e3, T, T / dt))
    res[prec]["frames_u8"] = h_out.clone()
if "--no-oracle" not in sys.argv:
    from oracle import eamm_oracle as oracle, kp_glue
    torch.set_num_threads(min(64, os.cpu_count()))
    t0 = time.perf_counter()
    o_deco = oracle.at_net2_forward(atsd, img, mfcc, pose, 1.6)
    o_src = oracle.kp_detector_forward(ksd, kcfg, img)
    o_drv = oracle.kp_detector_a_forward(asd, acfg, o_deco[0])
    o_init = {k: o_drv[k][:1] for k in ("value", "jacobian")}
    nv, nj = kp_glue.clip_glue(o_drv["value"], o_drv["jacobian"], None, None, o_src, o_init,
                               movement_scale=clip.movement_scale(o_src, o_init))
    frames = []
    for t in range(0, T, 16):
        n = min(16, T - t)
        o = oracle.generator_forward(sd, cfg, img.expand(n, -1, -1, -1).contiguous(), {"value": nv[t:t + n], "jacobian": nj[t:t + n]},
                                     {k: o_src[k].expand(n, *o_src[k].shape[1:]).contiguous() for k in ("value", "jacobian")})
        frames.append(oracle.frames_u8(o["prediction"]))
    want = torch.cat(frames).float()
    dt = time.perf_counter() - t0
    res["cpu_oracle"] = {"s_per_clip": dt, "frames_per_s": T / dt, "threads": torch.get_num_threads()}
    print("CPU oracle chain: %.1f s per clip, %.2f frames/s (%d threads)" % (dt, T / dt, torch.get_num_threads()))
    for prec in ("fp32", "bf16"):
        err = res[prec]["frames_u8"].float() - want
        mse = float((err ** 2).mean())
        res[prec]["psnr_db_u8"] = float(10 * np.log10(255.0 ** 2 / mse)) if mse > 0 else float("inf")
        res[prec]["max_abs_u8"] = int(err.abs().max())
        res[prec]["frac_exact_u8"] = float((err == 0).float().mean())
        print("%s: PSNR %.1f dB, max |diff| %d grey levels, %.2f%% of bytes identical" % (
            prec, res[prec]["psnr_db_u8"], res[prec]["max_abs_u8"], 100 * res[prec]["frac_exact_u8"]))
for prec in ("fp32", "bf16"):
    res[prec].pop("frames_u8", None)
if "--out" in sys.argv:
    with open(sys.argv[sys.argv.index("--out") + 1], "w") as f:
        json.dump(res, f, indent=1)
