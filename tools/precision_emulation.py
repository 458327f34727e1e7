"""tools/precision_emulation.py -- CPU emulation of operand-splitting schemes for the conv path: end-to-end error of the generator vs the fp32 oracle.
Products are formed from the rounded operands and accumulated in fp32 by F.conv2d (the tensor core's accumulate
truncation is NOT modelled: this isolates the operand-precision part of the error)."""
import sys, math, torch
import torch.nn.functional as F
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eamm_b200 import get_config, synth
from oracle import eamm_oracle as oracle

real_conv = F.conv2d
E4 = torch.float8_e4m3fn

def f8(x, scale):
    return (x * scale).clamp(-448, 448).to(E4).float() / scale

def pow2_scale(x):          # per-tensor power of two so that max|x| lands near 256
    m = x.abs().max().item()
    return 1.0 if m == 0 else 2.0 ** math.floor(math.log2(256.0 / m))

def mx8(x, dim):
    """MX-style e4m3: one power-of-two scale per block of 32 elements along `dim` (UE8M0 scales of
    tcgen05.mma.kind::mxf8f6f4.block_scale), block maximum mapped into [128, 256)."""
    n = x.shape[dim]
    pad = (-n) % 32
    xm = x.movedim(dim, -1)
    if pad:
        xm = F.pad(xm, (0, pad))
    blk = xm.reshape(*xm.shape[:-1], -1, 32)
    m = blk.abs().amax(-1, keepdim=True).clamp_min(2.0 ** -60)
    scale = torch.exp2(torch.floor(torch.log2(256.0 / m)))
    q = (blk * scale).clamp(-448, 448).to(E4).float() / scale
    q = q.reshape(*xm.shape)[..., :n]
    return q.movedim(-1, dim)


def make(scheme):
    def conv(x, weight=None, bias=None, padding=0, groups=1, **kw):
        if groups != 1:
            return real_conv(x, weight, bias, padding=padding, groups=groups, **kw)
        w = weight
        if scheme == "bf16x3":
            xh = x.bfloat16().float(); xl = (x - xh).bfloat16().float()
            wh = w.bfloat16().float(); wl = (w - wh).bfloat16().float()
            y = real_conv(xh, wl, None, padding=padding) + real_conv(xl, wh, None, padding=padding) + real_conv(xh, wh, None, padding=padding)
        elif scheme == "bf16":
            y = real_conv(x.bfloat16().float(), w.bfloat16().float(), None, padding=padding)
        elif scheme == "fp16":
            y = real_conv(x.half().float(), w.half().float(), None, padding=padding)
        elif scheme in ("f16+f8", "f16+f8_nohi8"):
            xh = x.half().float(); xr = x - xh
            wh = w.half().float(); wr = w - wh
            xl = f8(xr, pow2_scale(xr)); wl = f8(wr, pow2_scale(wr))
            if scheme == "f16+f8":
                x8 = f8(xh, pow2_scale(xh)); w8 = f8(wh, pow2_scale(wh))
            else:
                x8, w8 = xh, wh
            y = real_conv(x8, wl, None, padding=padding) + real_conv(xl, w8, None, padding=padding) + real_conv(xh, wh, None, padding=padding)
        elif scheme == "f16s+f8":
            # the pre-scaled domain of DESIGN.md section 9: max|A'| ~ 2^13, lo8 = e4m3((A' - hi) * 2^6), hi8 = e4m3(hi * 2^-6)
            def split(t):
                m = t.abs().max().item()
                sc = 1.0 if m == 0 else 2.0 ** math.floor(math.log2(2.0 ** 13 / m))
                tp = t * sc
                hi = tp.clamp(-65504, 65504).half().float()
                lo8 = ((tp - hi) * 64.0).clamp(-448, 448).to(E4).float()
                hi8 = (hi / 64.0).clamp(-448, 448).to(E4).float()
                return hi, lo8, hi8, sc
            xh, xl, x8, sx = split(x)
            wh, wl, w8, sw = split(w)
            y = (real_conv(x8, wl, None, padding=padding) + real_conv(xl, w8, None, padding=padding) +
                 real_conv(xh, wh, None, padding=padding)) / (sx * sw)
        elif scheme == "f16+mxf8":
            xh = x.half().float(); wh = w.half().float()
            xl = mx8(x - xh, 1); wl = mx8(w - wh, 1)
            x8 = mx8(xh, 1); w8 = mx8(wh, 1)
            y = real_conv(x8, wl, None, padding=padding) + real_conv(xl, w8, None, padding=padding) + real_conv(xh, wh, None, padding=padding)
        elif scheme == "bf16+f8":
            xh = x.bfloat16().float(); xr = x - xh
            wh = w.bfloat16().float(); wr = w - wh
            xl = f8(xr, pow2_scale(xr)); wl = f8(wr, pow2_scale(wr))
            x8 = f8(xh, pow2_scale(xh)); w8 = f8(wh, pow2_scale(wh))
            y = real_conv(x8, wl, None, padding=padding) + real_conv(xl, w8, None, padding=padding) + real_conv(xh, wh, None, padding=padding)
        else:
            raise ValueError(scheme)
        if bias is not None:
            y = y + bias.view(1, -1, 1, 1)
        return y
    return conv

cfg = get_config("full")
sd = synth.make_state_dict(cfg, seed=0)
B = int(os.environ.get("EMU_BATCH", "2"))
src, kpd, kps = synth.make_inputs(B, cfg, size=256, seed=1, shared_source=bool(int(os.environ.get("EMU_SHARED", "0"))))
torch.set_num_threads(8)
want = oracle.generator_forward(sd, cfg, src, kpd, kps)
for scheme in sys.argv[1:]:
    oracle.F.conv2d = make(scheme)
    try:
        got = oracle.generator_forward(sd, cfg, src, kpd, kps)
    finally:
        oracle.F.conv2d = real_conv
    print("%-14s" % scheme, " ".join("%s %.2e" % (k, (got[k] - want[k]).abs().max().item()) for k in ("prediction", "mask", "occlusion_map", "deformed")), flush=True)
