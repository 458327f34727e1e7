#!/usr/bin/env python
"""Stage-by-stage comparison of the CUDA path with the CPU oracle (run on the GPU box).

Prints max-abs / mean-abs error per intermediate tensor and per output; never raises on a numeric
mismatch so that one gpurun call reports everything.  Usage:
    python tools/gpu_stage_check.py [--config tiny|full] [--batch 2] [--precision fp32_simt]
"""
import argparse
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eamm_b200 import get_config, synth                                   # noqa: E402
from eamm_b200.modules.generator import OcclusionAwareGenerator           # noqa: E402
from oracle import eamm_oracle as oracle                                  # noqa: E402


def err(name, got, want):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    if got.shape != want.shape:
        print("%-28s SHAPE MISMATCH got %s want %s" % (name, tuple(got.shape), tuple(want.shape)))
        return
    d = (got - want).abs()
    bad = int((~torch.isfinite(got)).sum())
    print("%-28s max_abs %.3e  mean_abs %.3e  ref_absmax %.3e  nonfinite %d" %
          (name, d.max().item(), d.mean().item(), want.abs().max().item(), bad))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="tiny")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--precision", default="fp32_simt")
    ap.add_argument("--shared", action="store_true")
    args = ap.parse_args()
    cfg = get_config(args.config)
    size = args.size or (64 if args.config == "tiny" else 256)
    sd = synth.make_state_dict(cfg, seed=0)
    src, kpd, kps = synth.make_inputs(args.batch, cfg, size=size, seed=1, shared_source=args.shared)
    taps = {}
    t0 = time.time()
    want = oracle.generator_forward(sd, cfg, src, kpd, kps, taps=taps)
    print("oracle %.2fs" % (time.time() - t0))
    dev = torch.device("cuda:0")
    gen = OcclusionAwareGenerator(**cfg).eval()
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(dev)
    gen.precision = args.precision
    cu = lambda d: {k: v.to(dev) for k, v in d.items()}
    src_d = src.to(dev)
    if args.shared:
        src_d = src_d[:1].expand(args.batch, -1, -1, -1)
    try:
        got = gen(src_d, kp_driving=cu(kpd), kp_source=cu(kps))
        torch.cuda.synchronize()
    except Exception:
        traceback.print_exc()
        return 1
    eng = gen._eng
    B = args.batch
    h = size // eng.dm.step
    dws = eng.dm.ws[(B, size, size)]
    gws = eng.ws[(B, size, size)]
    nsrc = 1 if args.shared else B
    K1 = cfg["num_kp"] + 1
    print("---- config=%s batch=%d precision=%s shared=%s" % (args.config, B, args.precision, args.shared))
    nbk = cfg["num_bottleneck_blocks"]
    nbk_zero = nbk == 0
    err("source_small", dws.small[:nsrc, ..., :3].permute(0, 3, 1, 2), taps["source_small"][:nsrc])
    cat0 = dws.cat[0]
    err("hourglass_in", cat0.to_float(cat0.s_up, 4 * K1), taps["hourglass_in"])
    c_up = eng.dm.dec_ch[-1]
    err("hourglass_out[up]", cat0.to_float(0, c_up), taps["hourglass_out"][:, :c_up])
    err("first", gws.enc[0].to_float(0, cfg["block_expansion"])[:nsrc], taps["first"][:nsrc])
    enc_c = taps["encoded"].shape[1]
    err("encoded", gws.enc[-1].to_float(0, enc_c)[:nsrc], taps["encoded"][:nsrc])
    err("deformation", eng.last_dm["deformation"], taps["deformation"])
    err("warped", gws.x[0].to_float(0, enc_c), taps["warped"]) if nbk_zero else None
    cur = nbk % 2
    err("bottleneck", gws.x[cur].to_float(0, enc_c), taps["bottleneck"])
    err("decoded", gws.dec[-1].to_float(0, cfg["block_expansion"]), taps["decoded"])
    for k in ["mask", "sparse_deformed", "occlusion_map", "deformed", "prediction"]:
        err("out." + k, got[k], want[k])
    return 0


if __name__ == "__main__":
    sys.exit(main())
