#!/usr/bin/env python
"""Error of the fp32-equivalent modes against the CPU oracle, by scope of the mixed fp16 + fp8 scheme (GPU box tool).

    python tools/gpu_mix_error.py            # table: max-abs error per output for each variant and input size

Variants are selected through the environment before the engine is built (EAMM_B200_MIX = 0 | res | 1, EAMM_B200_MIX64,
EAMM_B200_MIX_SKIP = comma list of layer-name prefixes that stay on the 3-pass bf16 scheme)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eamm_b200 import get_config, synth                                   # noqa: E402
from eamm_b200.modules.generator import OcclusionAwareGenerator            # noqa: E402
from oracle import eamm_oracle as oracle                                   # noqa: E402  (checker)

VARIANTS = [("bf16x3 (MIX=0)", {"EAMM_B200_MIX": "0"}),
            ("bottleneck only (MIX=res)", {"EAMM_B200_MIX": "res"}),
            ("all eligible layers", {"EAMM_B200_MIX": "1", "EAMM_B200_MIX_HG": "1"}),
            ("all but hourglass (default)", {"EAMM_B200_MIX": "1"}),
            ("all but decoder", {"EAMM_B200_MIX": "1", "EAMM_B200_MIX_SKIP": "up"}),
            ("all but encoder", {"EAMM_B200_MIX": "1", "EAMM_B200_MIX_SKIP": "down,enc0", "EAMM_B200_MIX64": "0"})]
CASES = [(256, 8, 7), (512, 1, 562), (512, 2, 11), (128, 3, 178)]
KEYS = ("prediction", "mask", "occlusion_map", "deformed")


def main():
    dev = torch.device("cuda:0")
    cfg = get_config("full")
    sd = synth.make_state_dict(cfg, seed=0)
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    inputs, wants = {}, {}
    for size, batch, seed in CASES:
        inputs[(size, batch)] = synth.make_inputs(batch, cfg, size=size, seed=seed)
        wants[(size, batch)] = oracle.generator_forward(sd, cfg, *inputs[(size, batch)])
    print("%-28s %-12s" % ("variant", "case") + "".join("%14s" % k for k in KEYS))
    for name, env in VARIANTS:
        for k in ("EAMM_B200_MIX", "EAMM_B200_MIX64", "EAMM_B200_MIX_SKIP", "EAMM_B200_MIX_HG"):
            os.environ.pop(k, None)
        os.environ.update(env)
        gen = OcclusionAwareGenerator(**cfg).eval()
        gen.load_state_dict(sd)
        gen = gen.to(dev)
        gen.precision = "fp32"
        for (size, batch), (src, kpd, kps) in inputs.items():
            gen._eng = None                       # calibrate on this very input (what a first call does)
            out = gen(src.to(dev), kp_driving={k: v.to(dev) for k, v in kpd.items()},
                      kp_source={k: v.to(dev) for k, v in kps.items()})
            torch.cuda.synchronize()
            errs = [(out[k].cpu() - wants[(size, batch)][k]).abs().max().item() for k in KEYS]
            print("%-28s %-12s" % (name, "%dpx B=%d" % (size, batch)) + "".join("%14.3e" % e for e in errs), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
