#!/usr/bin/env python
"""Per-row roofline table from a committed bench line.

Re-presents MEASURED data only: the per-kernel CUDA-event times `bench.py --all-kernels` records in
`kernels_ms_per_step` / `hbm_kernels`, the algorithmic FLOPs of the reference convolutions (2*MAC, the layer table of
eamm_b200.synth.conv_layers = BASELINE.md section 2) and the measured peaks the bench line itself was scored against.
No GPU needed:

    python tools/roofline_table.py profiles/r2_bench_fp32_b32.json > profiles/r2_roofline_by_row.md
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eamm_b200 import get_config, synth          # noqa: E402


def layer_gflop_per_frame():
    """kernel name in the bench line -> (section-8 row, GFLOP per frame of the reference conv(s) it replaces)."""
    cfg = get_config("full")
    out = {}
    for prefix, cin, cout, k, kind in synth.conv_layers(cfg):
        if "hourglass.encoder" in prefix:
            i = int(prefix[-1]); hw, name, row = 64 >> i, "conv:hg.enc%d" % i, "a7"
        elif "hourglass.decoder" in prefix:
            i = int(prefix[-1]); hw, name, row = 4 << i, "conv:hg.dec%d" % i, "a7"
        elif prefix.startswith("dense_motion"):
            hw, name, row = 64, "conv:mask_occ", "a8"
        elif prefix.startswith("down_blocks"):
            i = int(prefix[-1]); hw, name, row = 256 >> i, "conv:down%d" % i, "a2"
        elif prefix.startswith("up_blocks"):
            i = int(prefix[-1]); hw, name, row = 128 << i, "conv:up%d" % i, "a11"
        elif prefix.startswith("bottleneck"):
            hw, name, row = 64, None, "a10"
        else:
            hw, name, row = 256, "conv:" + prefix, "a1" if prefix == "first" else "a12"
        gf = 2.0 * cin * cout * k * k * hw * hw / 1e9
        if name is None:                               # ResBlock2d: conv1 and conv2 are separate launches
            i = int(prefix[-1])
            out["conv:res%d.conv1" % i] = (row, gf)
            out["conv:res%d.conv2" % i] = (row, gf)
        else:
            r, g = out.get(name, (row, 0.0))
            out[name] = (r, g + gf)                     # mask + occlusion run as one merged conv
    return out


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_bench_fp32_b32.json")
    line = [json.loads(l) for l in open(path) if l.startswith("{")][0]
    B = line["config"]["global_batch"] // line["n_gpus"]
    prec = line["config"]["precision"]
    peak_tf = line["roofline"]["peak"]
    peak_gbs = line["roofline_hbm"]["peak"]
    burst_tf = peak_tf
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        burst_tf = json.load(open(mp)).get("bf16_tflops", peak_tf)
    kern = line["kernels_ms_per_step"]
    gflop = layer_gflop_per_frame()
    mixed = "e4m3" in line["dtype"]
    # bf16-pass equivalents the tensor pipe executes per algorithmic FLOP
    def passes(name):
        if prec not in ("fp32", "fp32_bf16x3"):
            return 1
        if mixed and name.startswith(("conv:down", "conv:res", "conv:up")):
            return 2
        return 3
    print("# Roofline by SURVEY §8(a) row — %s" % os.path.basename(path))
    print()
    print("Derived by `tools/roofline_table.py` from the bench line (B = %d frames, precision `%s`, %.3f ms per step, %s, "
          "SM clock %s MHz under `%s`): per-kernel CUDA-event times of one profiled step, algorithmic work of the reference "
          "ops, peaks = the ones the line was scored against (%.1f TFLOP/s dense bf16 sustained, %.1f GB/s HBM copy).  "
          "`executed` = algorithmic x the bf16-pass equivalents of the operand scheme (2 = fp16 + 2 x e4m3, 3 = bf16 hi/lo)."
          % (B, prec, line["ms_per_step"], line["unit"], line["clocks"]["sm_mhz"], ",".join(line["clocks"]["reasons"]) or "no throttle",
             peak_tf, peak_gbs))
    print()
    print("UP2 layers (UpBlock2d, hourglass decoder) execute 1/2.25 of the reference's MACs (nearest x2 + 3x3 run as four 2x2 "
          "convs on the low-res map), so their algorithmic rate can exceed the peak; `executed frac` = algorithmic / MAC "
          "reduction x passes / peak is the tensor-pipe view (values slightly above 1 = above the SUSTAINED cuBLAS figure, "
          "below the %.1f TFLOP/s burst figure of MEASURED_PEAKS.json)." % burst_tf)
    print()
    print("| row | kernel launch | ms | share of step | GFLOP (alg., x%d frames) | TFLOP/s alg. | frac of peak | MAC reduction | passes | executed frac |" % B)
    print("|---|---|---|---|---|---|---|---|---|---|")
    total_ms = sum(v["ms"] for v in kern.values())
    rows = {}
    for name, v in kern.items():
        if not name.startswith("conv:"):
            continue
        row, gf = gflop[name]
        tf = gf * B / v["ms"]                          # GFLOP / ms = TFLOP/s
        p = passes(name)
        red = 2.25 if name.startswith(("conv:up", "conv:hg.dec")) else 1.0
        rows.setdefault(row, [0.0, 0.0])
        rows[row][0] += v["ms"]; rows[row][1] += gf * B
        # 7x7 layers (3 or 12 real channels on one side) run as padded GEMMs (K 147 -> 448 for `first`, N 3 -> 28 per output
        # row for `final`, 12 -> 16 x 7 kx for mask+occlusion): their tensor-pipe work is not algorithmic x passes
        padded = name in ("conv:first", "conv:final", "conv:mask_occ")
        print("| %s | `%s` | %.4f | %.1f %% | %.1f | %.0f | %.2f | %s | %d | %s |"
              % (row, name[5:], v["ms"], 100 * v["ms"] / total_ms, gf * B, tf, tf / peak_tf, "2.25x" if red > 1 else "-", p,
                 "padded GEMM" if padded else "%.2f" % (tf / red * p / peak_tf)))
    print()
    print("| row (all its launches) | ms | share of step | TFLOP/s alg. | frac of peak |")
    print("|---|---|---|---|---|")
    for row in sorted(rows, key=lambda r: int(r[1:])):
        ms, gf = rows[row]
        print("| %s | %.3f | %.1f %% | %.0f | %.2f |" % (row, ms, 100 * ms / total_ms, gf / ms, gf / ms / peak_tf))
    conv_ms = sum(r[0] for r in rows.values()); conv_gf = sum(r[1] for r in rows.values())
    print("| all convs | %.3f | %.1f %% | %.0f | %.2f |" % (conv_ms, 100 * conv_ms / total_ms, conv_gf / conv_ms,
                                                         conv_gf / conv_ms / peak_tf))
    print()
    print("HBM-bound kernels (algorithmic bytes per launch as in SURVEY §8(d), fp32-equivalent storage):")
    print()
    print("| row | kernel | ms | share of step | MB (alg.) | GB/s | frac of HBM copy peak |")
    print("|---|---|---|---|---|---|---|")
    hrow = {"pack_image": "a1 (layout)", "aa_downsample": "a3", "kp_stage": "a4-a6", "flow_combine": "a8", "warp_occlude": "a9-i",
            "warp_image": "a9-ii", "nchw_to_act": "a1 (layout)"}
    for name, v in sorted(line["hbm_kernels"].items(), key=lambda kv: -kv[1]["ms"]):
        print("| %s | `%s` | %.4f | %.1f %% | %.2f | %.0f | %.2f |"
              % (hrow.get(name, ""), name, v["ms"], 100 * v["ms"] / total_ms, v["mbytes"], v["gbs"], v["gbs"] / peak_gbs))
    print()
    print("Sum of the per-kernel intervals %.3f ms (each bracketed by its own event pair) vs %.3f ms per step in the timed "
          "loop; a13 (eval BatchNorm) has no kernel: folded into weights / epilogues." % (total_ms, line["ms_per_step"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
