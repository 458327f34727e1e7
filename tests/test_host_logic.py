"""CPU tests: host-side packing logic, the numpy sampler pins, the drop-in module surface, the C ABI
exports and the multi-process sharding helpers (gloo, world_size 2).  No GPU compute here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from eamm_b200 import get_config, synth, sharding
from eamm_b200 import engine
from oracle import eamm_oracle as oracle
from oracle import sampler_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ weight packing
L_UP2 = 2          # _lib.CONV_UP2_3X3


def test_fold_bn_matches_conv_then_bn():
    g = torch.Generator().manual_seed(0)
    w, b = torch.randn(8, 4, 3, 3, generator=g), torch.randn(8, generator=g)
    bn = {"weight": torch.rand(8, generator=g) + 0.5, "bias": torch.randn(8, generator=g),
          "running_mean": torch.randn(8, generator=g), "running_var": torch.rand(8, generator=g) + 0.5}
    x = torch.randn(2, 4, 9, 9, generator=g)
    want = F.batch_norm(F.conv2d(x, w, b, padding=1), bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"],
                        False, 0.1, 1e-5)
    wf, bf = engine.fold_bn(w, b, bn)
    assert torch.allclose(F.conv2d(x, wf, bf, padding=1), want, atol=1e-5)
    s, t = engine.bn_affine(bn)
    want2 = F.batch_norm(x[:, :4].repeat(1, 2, 1, 1), bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"],
                         False, 0.1, 1e-5)
    assert torch.allclose(x.repeat(1, 2, 1, 1) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1), want2, atol=1e-5)


def test_up2_parity_weights_equal_upsample_then_conv():
    """nearest x2 + 3x3 pad 1 == four 2x2 convs on the low-res input (UpBlock2d, util.py:895-897)."""
    g = torch.Generator().manual_seed(1)
    w = torch.randn(5, 3, 3, 3, generator=g)
    x = torch.randn(2, 3, 6, 4, generator=g)
    want = F.conv2d(F.interpolate(x, scale_factor=2), w, padding=1)
    pw = engine.up2_parity_weights(w)                      # [4][4][cout][cin]
    got = torch.zeros_like(want)
    xp = F.pad(x, (1, 1, 1, 1))
    H, W = x.shape[2:]
    for a in (0, 1):
        for b in (0, 1):
            acc = 0
            for ty in (0, 1):
                for tx in (0, 1):
                    dy, dx = a - 1 + ty, b - 1 + tx
                    sl = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                    acc = acc + torch.einsum("oc,nchw->nohw", pw[a * 2 + b, ty * 2 + tx], sl)
            got[:, :, a::2, b::2] = acc
    assert torch.allclose(got, want, atol=1e-5)


def test_tc_weight_packing_order_and_split():
    g = torch.Generator().manual_seed(2)
    full = torch.randn(9, 16, 64, generator=g)             # [taps][cout][cin]
    p1 = engine.pack_tc_weights(full, 1, 1)
    assert p1.shape == (16, 9 * 64) and p1.dtype == torch.bfloat16
    assert torch.equal(p1[3, 2 * 64:3 * 64], full[2, 3].bfloat16())
    p3 = engine.pack_tc_weights(full, 1, 3)
    assert p3.shape == (16, 3 * 9 * 64)
    blk = lambda ps, t: p3[:, (ps * 9 + t) * 64:(ps * 9 + t + 1) * 64].float()      # K order (pass, tap, channel)
    lo, hi, hi2 = blk(0, 2), blk(1, 2), blk(2, 2)
    assert torch.equal(hi, hi2)
    assert (hi + lo - full[2]).abs().max() < 2e-5          # 16 mantissa bits survive the split
    up = engine.pack_tc_weights(torch.randn(16, 8, 64, generator=g), 4, 1)
    assert up.shape == (4 * 8, 4 * 64)


def test_row7_packed_first_conv_and_fold_packings_reproduce_the_conv():
    """`first` (generator.py:25) as EAMM_CONV_ROW7_PACKED: the source packed to 16-byte pixels [hi0,hi1,hi2,0,lo0,lo1,lo2,0] with a
    3-row / 3-column zero border, one 64-wide K window per filter row = 8 consecutive packed pixels (kx = 7 is padding); the
    weight matrices of engine.pack_tc_weights_row7 (one plane, hi/lo in two passes, hi/lo folded along N) times those windows
    equal the 7x7 / pad-3 convolution of the split operands.  Likewise fold scheme 1 (pack_tc_weights_fold) for a 3x3 layer."""
    g = torch.Generator().manual_seed(11)
    H, W, cout = 6, 9, 8
    x = torch.rand(3, H, W, generator=g)
    w = torch.randn(cout, 3, 7, 7, generator=g) * 0.1
    a_hi = x.bfloat16().float()
    a_lo = (x - a_hi).bfloat16().float()
    w_hi = w.bfloat16().float()
    w_lo = (w - w_hi).bfloat16().float()
    conv = lambda a, b: F.conv2d(a[None].double(), b.double(), padding=3)[0]
    packed = torch.zeros(H + 6, W + 8, 8, dtype=torch.float64)
    packed[3:3 + H, 3:3 + W, 0:3] = a_hi.permute(1, 2, 0).double()

    def windows():                                        # A [H*W][7 ky][64]: pixel (y, x), filter row ky -> packed[y+ky, x:x+8]
        return torch.stack([torch.stack([packed[y + ky, xx:xx + 8].reshape(64) for ky in range(7)])
                            for y in range(H) for xx in range(W)])

    # one plane (bf16 / fp16 modes): K = (ky, kx, channel)
    w1 = engine.pack_tc_weights_row7(w, 16, 1).double()
    assert w1.shape == (16, 7 * 64)
    D = windows().reshape(H * W, -1) @ w1.T
    got = D[:, :cout].T.reshape(cout, H, W)
    assert torch.allclose(got, conv(a_hi, w_hi), atol=1e-12) and D[:, cout:].abs().max() == 0
    # hi/lo planes, two passes: pass 0 = w_lo against the hi channels, pass 1 = w_hi against hi and lo channels
    packed[3:3 + H, 3:3 + W, 4:7] = a_lo.permute(1, 2, 0).double()
    want = conv(a_hi, w_lo) + conv(a_hi + a_lo, w_hi)
    assert (want - conv(x, w)).abs().max() <= 2e-5 * conv(x, w).abs().max()        # 16 mantissa bits per operand
    A = windows()
    w2 = engine.pack_tc_weights_row7(w, 16, 2).double().view(16, 2, 7 * 64)
    D = A.reshape(H * W, -1) @ w2[:, 0].T + A.reshape(H * W, -1) @ w2[:, 1].T
    assert torch.allclose(D[:, :cout].T.reshape(cout, H, W), want, atol=1e-12)
    # folded along N: rows [w_hi vs (a_hi, a_lo) | w_lo vs a_hi]; the epilogue adds the two halves
    wf = engine.pack_tc_weights_row7(w, 16, 2, fold=2).double()
    assert wf.shape == (32, 7 * 64)
    D = A.reshape(H * W, -1) @ wf.T
    assert torch.allclose((D[:, :16] + D[:, 16:])[:, :cout].T.reshape(cout, H, W), want, atol=1e-12)

    # fold scheme 1, 3x3: per tap a_lo x b_hi (N = bn) and a_hi x [b_hi; b_lo] (N = 2 bn), halves added in the epilogue
    cin, co, hh = 64, 16, 5
    xa = torch.randn(cin, hh, hh, generator=g)
    wa = torch.randn(co, cin, 3, 3, generator=g) * 0.05
    full = wa.permute(2, 3, 0, 1).reshape(9, co, cin)
    pf = engine.pack_tc_weights_fold(full, 1).double()                             # [2*co][9*cin]
    assert pf.shape == (2 * co, 9 * cin)
    xh = xa.bfloat16().float(); xl = (xa - xh).bfloat16().float()
    wh = wa.bfloat16().float(); wl = (wa - wh).bfloat16().float()
    c3 = lambda a, b: F.conv2d(a[None].double(), b.double(), padding=1)[0]
    want3 = c3(xl, wh) + c3(xh, wh) + c3(xh, wl)
    pad = lambda t: F.pad(t.double(), (1, 1, 1, 1))
    ph, pl = pad(xh), pad(xl)
    got3 = torch.zeros(co, hh, hh, dtype=torch.float64)
    for y in range(hh):
        for xx in range(hh):
            acc = torch.zeros(2 * co, dtype=torch.float64)
            for t in range(9):
                ky, kx = divmod(t, 3)
                blk = pf[:, t * cin:(t + 1) * cin]
                acc[:co] += blk[:co] @ pl[:, y + ky, xx + kx]                       # a_lo x b_hi
                acc += blk @ ph[:, y + ky, xx + kx]                                 # a_hi x [b_hi; b_lo]
            got3[:, y, xx] = acc[:co] + acc[co:]
    assert torch.allclose(got3, want3, atol=1e-10)


def test_mix_calibration_exponents_centre_the_maxima_with_hysteresis():
    """engine.MixCalib (fp16 + 2 x e4m3 activations): the pre-scale exponent puts a tensor's maximum at 2^11.5 (+-0.5 octave:
    4x below the e4m3 lo8 saturation, 22x below fp16 overflow); later forwards only move a tensor whose stored maximum left
    [2^9, 2^13.2]; tensors that were never written (0) or overflowed (inf / nan) keep their exponent."""
    import math
    cal = engine.MixCalib()
    for nm in ("a", "b", "c", "d", "e"):
        cal.register(nm)
    assert cal.register("a") == "a" and len(cal.slots) == 5          # idempotent
    cal.alloc(torch.device("cpu"))
    cal.begin(True)
    assert cal.track and cal.ptr("b") == cal.amax.data_ptr() + 4
    cal.amax.copy_(torch.tensor([3.7, 1e-3, 900.0, 0.0, float("nan")]))
    new = cal.proposal(force=True)
    for nm, v in (("a", 3.7), ("b", 1e-3), ("c", 900.0)):
        stored = math.log2(v) + new[nm]
        assert abs(stored - engine.ACT_TOP) <= 0.5 + 1e-9, (nm, stored)
    assert new["d"] == 0 and new["e"] == 0
    cal.exps = new
    # drift inside the window: nothing moves; outside: only that tensor is re-centred
    cal.amax.copy_(torch.tensor([3.7 * 2.0, 1e-3 / 3.0, 900.0 * 3.5, 0.0, 1.0]))
    again = cal.proposal(force=False)
    assert again["a"] == new["a"] and again["b"] == new["b"]
    assert again["c"] != new["c"] and abs(math.log2(900.0 * 3.5) + again["c"] - engine.ACT_TOP) <= 0.5 + 1e-9
    assert again["e"] != 0                                             # first real value of a late tensor
    cal.begin(False)
    assert not cal.track and cal.ptr("a") is None                      # untracked forwards pass no statistics pointer
    # weight side: one exponent per output channel puts the row maximum just below 2^13.4, lo8 within the e4m3 range
    g = torch.Generator().manual_seed(4)
    full = torch.randn(9, 32, 128, generator=g) * torch.logspace(-3, 1, 32).view(1, 32, 1)
    packed, e = engine.pack_tc_weights_mix(full, 1)
    assert packed.shape == (32, 9 * 128 * 4) and packed.dtype == torch.uint8 and e.dtype == torch.int32
    top = torch.log2(full.abs().amax(dim=(0, 2))) + e.float()
    assert (top <= engine.MIX_TOP).all() and (top > engine.MIX_TOP - 1.0).all()


def test_clip_input_preparation_matches_the_reference_loops():
    """eamm_b200.clip host-side input preparation against a literal restatement of demo.py:296-340 (`test_auido`): the MFCC
    window loop (:321-328) and the pose mirroring / tiling / cut (:297-299, :334-340), for utterance and pose lengths on
    both sides of every branch.  Index work: bit-exact."""
    from eamm_b200 import clip
    rng = np.random.default_rng(0)

    def ref_windows(mfcc):                                   # demo.py:321-328
        ind, out = 3, []
        while ind <= int(mfcc.shape[0] / 4) - 4:
            out.append(mfcc[(ind - 3) * 4: (ind + 4) * 4, 1:])
            ind += 1
        return out

    def ref_pose(all_pose, n_frames):                        # demo.py:297-299, 334-340
        pose = all_pose[:, :6]
        if len(pose) == 1:
            pose = np.repeat(pose, 100, 0)
        if len(pose) < n_frames:
            gap = n_frames - len(pose)
            n = int((gap / len(pose) / 2)) + 2
            pose = np.concatenate((pose, pose[::-1, :]), axis=0)
            pose = np.tile(pose, (n, 1))
        if len(pose) > n_frames:
            pose = pose[:n_frames, :]
        return pose

    for n in (0, 23, 27, 28, 31, 32, 100, 401, 1203):
        mfcc = rng.standard_normal((n, 13)).astype(np.float32)
        want = ref_windows(mfcc)
        got = clip.mfcc_windows(mfcc)
        assert got.shape == (len(want), 28, 12) and got.dtype == np.float32
        assert len(want) == max(0, n // 4 - 6)
        if want:
            assert np.array_equal(got, np.stack(want))
    with pytest.raises(ValueError):
        clip.mfcc_windows(np.zeros((40,), dtype=np.float32))
    for n_win, n_pose, T in ((30, 30, 300), (30, 1, 300), (30, 7, 64), (30, 90, 40), (12, 299, 300), (5, 300, 300), (9, 3, 9)):
        win13 = rng.standard_normal((n_win, 28, 13)).astype(np.float32)
        pose7 = rng.standard_normal((n_pose, 7)).astype(np.float32)
        m, p = clip.clip_inputs_from_windows(win13, pose7, T=T)
        assert tuple(m.shape) == (1, T, 28, 12) and tuple(p.shape) == (1, T, 6)
        assert np.array_equal(p[0].numpy(), ref_pose(pose7, T).astype(np.float32))
        for t in (0, T // 2, T - 1):
            assert np.array_equal(m[0, t].numpy(), win13[t % n_win, :, 1:])


def test_workspace_cache_lru_and_keypoint_struct_rules():
    """Host plumbing of engine.py that needs no GPU: the bounded per-shape workspace cache (least recently USED shape is
    dropped) and the eamm_kp struct built from the caller's dicts (dense_motion.py:55: 'jacobian' is optional; a leading
    dimension of 1 or an expanded stride-0 batch broadcasts; anything else is refused with the reference's shapes named)."""
    ws = engine.WorkspaceCache(limit=2)
    ws[(1, 256, 256)] = "a"
    ws[(32, 256, 256)] = "b"
    assert ws.get((1, 256, 256)) == "a"                      # touch: (32, ...) is now the oldest
    ws[(8, 256, 256)] = "c"
    assert (32, 256, 256) not in ws and (1, 256, 256) in ws and len(ws) == 2 and list(ws.values()) == ["a", "c"]
    assert ws.get((5, 5, 5)) is None
    ws.clear()
    assert len(ws) == 0

    cpu = torch.device("cpu")
    B, K = 4, 10
    kp = {"value": torch.rand(B, K, 2), "jacobian": torch.rand(B, K, 2, 2), "heatmap": torch.rand(B, K, 58, 58)}   # extra key ignored
    s, keep = engine._kp_struct(kp, B, K, cpu)
    assert (s.value, s.jacobian) == (kp["value"].data_ptr(), kp["jacobian"].data_ptr())
    assert (s.value_stride, s.jacobian_stride) == (K * 2, K * 4) and len(keep) == 2
    s, _ = engine._kp_struct({"value": kp["value"]}, B, K, cpu)                       # no Jacobians
    assert s.jacobian is None
    one = {"value": torch.rand(1, K, 2), "jacobian": torch.rand(1, K, 2, 2)}
    for d in (one, {k: v.expand(B, *v.shape[1:]) for k, v in one.items()}):             # batch 1 / expanded: broadcast
        s, keep = engine._kp_struct(d, B, K, cpu)
        assert (s.value_stride, s.jacobian_stride) == (0, 0) and keep[0].shape[0] == 1
    nc = kp["value"].transpose(1, 2).contiguous().transpose(1, 2)                        # right shape, wrong strides: copied
    s, keep = engine._kp_struct({"value": nc}, B, K, cpu)
    assert keep[0].is_contiguous() and torch.equal(keep[0], nc)
    for bad in ({"value": torch.rand(B, K, 3)}, {"value": torch.rand(3, K, 2)}, {"value": torch.rand(B, K, 2).double()},
                {"value": kp["value"], "jacobian": torch.rand(B, K, 4)}, {"value": kp["value"].numpy()}):
        with pytest.raises(RuntimeError):
            engine._kp_struct(bad, B, K, cpu)


def test_anti_alias_kernel_is_separable_and_the_subsample_phase_is_zero():
    """a3 (util.py:1005-1052): the engines hand eamm_aa_downsample the 1-D factor g1 = rowsum(k2) / sum of the reference's
    13x13 Gaussian buffer (engine.py DenseMotionEngine._pack, kp_engine.py) and the kernel filters rows, then columns.  Pins
    the two facts that rewrite rests on: k2 is exactly the outer product of that factor, and pad 6 -> depthwise conv -> `::4`
    keeps output pixel (i, j) centred on input pixel (4i, 4j) -- phase 0, integer-exact."""
    w = synth.aa_kernel(3)                                   # [3,1,13,13], the buffer `down.weight` of the reference
    k2 = w[0, 0].double()
    assert w.shape == (3, 1, 13, 13) and abs(k2.sum().item() - 1.0) < 1e-6 and torch.equal(w[0], w[1]) and torch.equal(w[1], w[2])
    g1 = k2.sum(1)
    g1 = g1 / g1.sum()
    assert (torch.outer(g1, g1) - k2).abs().max().item() < 1e-8
    assert torch.allclose(g1, g1.flip(0), atol=1e-12)        # symmetric: correlation == convolution
    g = torch.Generator().manual_seed(9)
    x = torch.rand(1, 3, 64, 64, generator=g, dtype=torch.float64)
    want = F.conv2d(F.pad(x, (6, 6, 6, 6)), w.double(), groups=3)[:, :, ::4, ::4]        # util.py:1048-1050
    xp = F.pad(x, (6, 6, 6, 6))
    rows = sum(g1[t] * xp[:, :, :, t:t + 64] for t in range(13))                          # filter along x, all padded rows
    got = sum(g1[t] * rows[:, :, t:t + 64, :] for t in range(13))[:, :, ::4, ::4]         # then along y, subsample
    assert got.shape == (1, 3, 16, 16) and (got - want).abs().max().item() < 5e-8      # k2 is the fp32 rounding of the outer product
    # phase: an impulse at input (4i, 4j) lands with the kernel's centre weight on output (i, j) and nowhere stronger
    imp = torch.zeros(1, 3, 64, 64, dtype=torch.float64)
    imp[0, :, 20, 36] = 1.0
    out = F.conv2d(F.pad(imp, (6, 6, 6, 6)), w.double(), groups=3)[:, :, ::4, ::4]
    assert out[0, 0].argmax().item() == 5 * 16 + 9 and abs(out[0, 0, 5, 9].item() - k2[6, 6].item()) < 1e-15


# ------------------------------------------------------------------ sampler pins (index selection)
def _emulate_kxn_tile_gemm(xpad, pad, y_rows, x_cols, B, cin, ntap):
    """D[128, 112] of one conv_tc tile: K loop over `ntap` input-row taps; A rows = the listed (y, x) pixels."""
    D = torch.zeros(len(y_rows), B.shape[0], dtype=torch.float64)
    for j in range(ntap):
        zero = torch.zeros(cin, dtype=torch.float64)                 # TMA zero-fills out-of-range pixels
        A = torch.stack([xpad[y + j - 3 + pad, x + pad] if 0 <= x + pad < xpad.shape[1] and
                         0 <= y + j - 3 + pad < xpad.shape[0] else zero for y, x in zip(y_rows, x_cols)])    # [128][cin]
        D += A @ B[:, j * cin:(j + 1) * cin].T
    return D


def test_kxn_wide_packings_reproduce_the_7x7_conv():
    """Schemes 3 and 4 of eamm_conv_tc (112-column kx-in-N): the packed weight matrices plus the shift-sum
    of the kernel's epilogue (conv_tc.cu epilogue_kxn_wide) equal a 7x7 / pad-3 convolution."""
    g = torch.Generator().manual_seed(5)
    cin, pad = 64, 16
    # scheme 3: 3 NCHW channels, four output rows per tile, x tiles of 122 (+6 halo) pixels
    H, W = 8, 130
    w = torch.randn(3, cin, 7, 7, generator=g, dtype=torch.float64)
    x = torch.randn(H, W, cin, generator=g, dtype=torch.float64)
    want = F.conv2d(x.permute(2, 0, 1)[None], w, padding=3)[0]                     # [3][H][W]
    full = torch.zeros(49, 16, cin, dtype=torch.float64)
    full[:, :3] = w.permute(2, 3, 0, 1).reshape(49, 3, cin)
    xpad = F.pad(x, (0, 0, pad, pad, pad, pad))
    rows = torch.zeros(4, 7, 4, 10, cin, dtype=torch.float64)
    for dr in range(4):
        rows[dr, :, :3, dr:dr + 7] = full.view(7, 7, 16, cin)[:, :, :3].permute(1, 2, 0, 3)
    B = rows.reshape(112, 10 * cin)
    packed = engine.pack_tc_weights_kxn_rows(full.float(), 3, 1, 0).double()          # bf16-rounded copy of B
    assert packed.shape == B.shape and (packed - B).abs().max() <= B.abs().max() * 2 ** -8
    folded = engine.pack_tc_weights_kxn_rows(full.float(), 3, 3, 1).double()
    assert folded.shape == (224, 10 * cin)
    assert (folded[:112] + folded[112:] - B).abs().max() <= B.abs().max() * 2 ** -15
    split3 = engine.pack_tc_weights_kxn_rows(full.float(), 3, 3, 0).double()
    assert split3.shape == (112, 30 * cin) and torch.equal(split3[:, 10 * cin:20 * cin], folded[:112])
    got = torch.zeros_like(want)
    for y0 in range(0, H, 4):
        for x0 in range(0, W, 122):
            D = _emulate_kxn_tile_gemm(xpad, pad, [y0] * 128, [x0 - 3 + p for p in range(128)], B, cin, 10)
            for dr in range(4):
                for r in range(122):
                    if x0 + r < W and y0 + dr < H:
                        for co in range(3):
                            got[co, y0 + dr, x0 + r] = sum(D[r + kx, dr * 28 + kx * 4 + co] for kx in range(7))
    assert torch.allclose(got, want, atol=1e-9)
    # scheme 4: 16 fp32 NHWC couts, W = 64, tile = two whole rows, zero padding instead of an x halo
    H, W = 4, 64
    w = torch.randn(16, cin, 7, 7, generator=g, dtype=torch.float64)
    x = torch.randn(H, W, cin, generator=g, dtype=torch.float64)
    want = F.conv2d(x.permute(2, 0, 1)[None], w, padding=3)[0]
    full = w.permute(2, 3, 0, 1).reshape(49, 16, cin)
    B = full.view(7, 7, 16, cin).permute(1, 2, 0, 3).reshape(112, 7 * cin)
    packed = engine.pack_tc_weights_kxn_full(full.float(), 1, 0).double()
    assert packed.shape == B.shape and (packed - B).abs().max() <= B.abs().max() * 2 ** -8
    xpad = F.pad(x, (0, 0, pad, pad, pad, pad))
    got = torch.zeros_like(want)
    for y0 in range(0, H, 2):
        px = [(y0 + p // W, p % W) for p in range(128)]
        D = _emulate_kxn_tile_gemm(xpad, pad, [a for a, _ in px], [b for _, b in px], B, cin, 7)
        for p in range(128):
            yl, xl = p // W, p % W
            for co in range(16):
                got[co, y0 + yl, xl] = sum(D[p + kx - 3, kx * 16 + co] for kx in range(7) if 0 <= xl + kx - 3 < W)
    assert torch.allclose(got, want, atol=1e-9)


def _plan(lib, L, kind, cin, cout, n, h, w, planes, flags=0, nhwc=False, nchw_c=0, f16=False):
    """eamm_conv_tc_query on dummy pointers (nothing is dereferenced or launched by the dry run).
    f16: EAMM_F16 operands (planes 1 = fp16, 2 = mixed fp16 + 2 x e4m3)."""
    a = L.ConvArgs()
    a.kind, a.flags, a.cin, a.cout = kind, flags, cin, cout
    dt = L.EAMM_F16 if f16 else L.EAMM_BF16
    act = L.Act(data=4096, dtype=dt, n=n, h=h, w=w, c=cin, c_off=0, c_buf=cin, planes=planes,
                n_stride=h * w * planes * cin)
    oh, ow = (2 * h, 2 * w) if kind == L.CONV_UP2_3X3 else ((h // 2, w // 2) if flags & L.EPI_POOL2 else (h, w))
    out = L.Act(data=4096, dtype=dt, n=n, h=oh, w=ow, c=cout, c_off=0, c_buf=cout, planes=planes,
                n_stride=oh * ow * planes * cout)
    a.inp, a.weight, a.bias = ctypes.pointer(act), 4096, 4096
    if nhwc:
        a.out_nhwc_f32 = 4096
    elif nchw_c:
        a.out_nchw, a.out_nchw_c = 4096, nchw_c
    else:
        a.out = ctypes.pointer(out)
    a.splitk_ws, a.splitk_ws_bytes = 4096, L.SPLITK_WS_BYTES
    q = (ctypes.c_int * 6)()
    assert lib.eamm_conv_tc_query(ctypes.byref(a), q) == 0
    return {"bn": q[0], "scheme": q[1], "fold": q[2], "ksub": q[3], "pair": q[4] & 1, "halo_tile": (q[4] >> 1) & 1,
            "splitk": q[4] >> 8, "stages": q[5]}


def test_conv_tc_planner_decisions_for_the_path_layers():
    """Planning dry run of eamm_conv_tc for a 148-SM device (EAMM_TC_NUM_SMS, no GPU needed): the schemes DESIGN.md
    describes are the ones chosen for the BASELINE configs[1] layers and for the batch-1 per-frame path."""
    code = r"""
import ctypes, json, os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
os.environ["EAMM_TC_NUM_SMS"] = "148"
for k in list(os.environ):
    if k.startswith("EAMM_TC_") and k != "EAMM_TC_NUM_SMS":
        del os.environ[k]
from eamm_b200 import _lib as L
from test_host_logic import _plan
lib = L.load()
P = lambda *a, **k: _plan(lib, L, *a, **k)
out = {}
for B in (1, 32):
    out["res%%d" %% B] = P(L.CONV_3X3, 256, 256, B, 64, 64, 2)
    out["res_bf16_%%d" %% B] = P(L.CONV_3X3, 256, 256, B, 64, 64, 1)
    out["down0_%%d" %% B] = P(L.CONV_3X3, 64, 128, B, 256, 256, 2, flags=3)
    out["enc4_%%d" %% B] = P(L.CONV_3X3, 1024, 1024, B, 4, 4, 2, flags=3)
    out["dec1_%%d" %% B] = P(L.CONV_UP2_3X3, 2048, 512, B, 4, 4, 2, flags=1)
    out["mask%%d" %% B] = P(L.CONV_7X7, 128, 16, B, 64, 64, 2, nhwc=True)
    out["final%%d" %% B] = P(L.CONV_7X7, 64, 16, B, 256, 256, 2, flags=4, nchw_c=3)
    out["final_bf16_%%d" %% B] = P(L.CONV_7X7, 64, 16, B, 256, 256, 1, flags=4, nchw_c=3)
    out["res_mix_%%d" %% B] = P(L.CONV_3X3, 256, 256, B, 64, 64, 2, f16=True)
    out["res_f16_%%d" %% B] = P(L.CONV_3X3, 256, 256, B, 64, 64, 1, f16=True)
    out["final_f16_%%d" %% B] = P(L.CONV_7X7, 64, 16, B, 256, 256, 1, flags=4, nchw_c=3, f16=True)
    out["down0_mix_%%d" %% B] = P(L.CONV_3X3, 64, 128, B, 256, 256, 2, flags=3, f16=True)
    out["up1_mix_%%d" %% B] = P(L.CONV_UP2_3X3, 128, 64, B, 128, 128, 2, flags=1, f16=True)
    out["up0_f16_%%d" %% B] = P(L.CONV_UP2_3X3, 256, 128, B, 64, 64, 1, flags=1, f16=True)
print(json.dumps(out))
""" % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    p = json.loads(r.stdout.strip().splitlines()[-1])
    # bottleneck convs at B=32: N = 256 CTA pairs, three passes unfolded, no split; at B=1 one wave of N = 64 tiles
    assert (p["res32"]["bn"], p["res32"]["pair"], p["res32"]["fold"], p["res32"]["splitk"]) == (256, 1, 0, 1)
    assert (p["res1"]["bn"], p["res1"]["pair"], p["res1"]["splitk"]) == (64, 0, 1)
    assert p["res_bf16_1"]["bn"] == 64 and p["res_bf16_32"]["bn"] == 256
    # cout <= 128 in split mode: weight planes folded into N, CTA pairs once the chip is filled
    assert (p["down0_32"]["bn"], p["down0_32"]["fold"], p["down0_32"]["pair"]) == (128, 1, 1)
    # small hourglass maps: N = 256 tiles with split-K (<= 9 splits, tiles * splits <= 148 SMs)
    # (batch 1: one M tile -- 64-column tiles so that 16 N tiles x 9 splits stream the weights through 144 SMs, not 36)
    assert (p["enc4_32"]["bn"], p["enc4_32"]["splitk"]) == (256, 9) and (p["enc4_1"]["bn"], p["enc4_1"]["splitk"]) == (64, 9)
    assert (p["dec1_32"]["bn"], p["dec1_32"]["splitk"]) == (256, 4)
    # 7x7 layers: 112-column kx-in-N schemes (4 = full-width mask+occlusion, 3 = four output rows for `final`), folded
    for B in (1, 32):
        assert (p["mask%d" % B]["scheme"], p["mask%d" % B]["bn"], p["mask%d" % B]["fold"]) == (4, 112, 1)
        assert (p["final%d" % B]["scheme"], p["final%d" % B]["bn"], p["final%d" % B]["fold"]) == (3, 112, 1)
        assert p["final%d" % B]["stages"] == 4                    # compact epilogue buffer -> fourth pipeline stage
        assert (p["final_bf16_%d" % B]["scheme"], p["final_bf16_%d" % B]["fold"]) == (3, 0)
        assert (p["final_f16_%d" % B]["scheme"], p["final_f16_%d" % B]["fold"]) == (3, 0)
    # mixed fp16 + fp8 bottleneck convs: N = 256 CTA pairs like the bf16 hi/lo scheme, never folded, no split at B = 32
    assert (p["res_mix_32"]["bn"], p["res_mix_32"]["pair"], p["res_mix_32"]["fold"], p["res_mix_32"]["splitk"]) == (256, 1, 0, 1)
    assert p["res_mix_1"]["fold"] == 0 and p["res_f16_32"]["bn"] == 256 and p["res_f16_32"]["pair"] == 1
    # halo-tile scheme (one 10x18 halo tile per K chunk, taps as descriptor views): every 3x3 / UP2 layer with fp16 or mixed
    # operands whose 8x16 tiles fill the chip -- the B=32 generator layers, not the batch-1 calls, never the bf16 hi/lo layers
    for k in ("res_mix_32", "res_f16_32", "down0_mix_32", "up1_mix_32", "up0_f16_32", "down0_mix_1"):
        assert p[k]["halo_tile"] == 1 and p[k]["pair"] == 1 and p[k]["fold"] == 0 and p[k]["splitk"] == 1, (k, p[k])
    # batch 1: the narrowest N tile when it still occupies >= 60 % of the SMs (single CTAs: pairs need a full chip)
    for k in ("res_mix_1", "res_f16_1", "up1_mix_1"):
        assert (p[k]["halo_tile"], p[k]["bn"], p[k]["pair"], p[k]["splitk"]) == (1, 64, 0, 1), (k, p[k])
    assert p["up0_f16_1"]["halo_tile"] == 0                     # 64 items for 148 SMs: the per-tap scheme with its own N tile
    for k in ("res32", "res1", "down0_32", "enc4_32", "mask32", "final32"):
        assert p[k]["halo_tile"] == 0, (k, p[k])
    # weight stage = a filter row for the narrow N tiles (3 taps), one tap at N = 256; UP2: a whole 2x2 class
    assert (p["down0_mix_32"]["bn"], p["down0_mix_32"]["ksub"]) == (128, 3) and p["res_mix_32"]["ksub"] == 1
    assert (p["up1_mix_32"]["bn"], p["up1_mix_32"]["ksub"]) == (64, 4) and (p["up0_f16_32"]["bn"], p["up0_f16_32"]["ksub"]) == (128, 4)


def test_conv_tc_planner_accepts_every_path_layer_at_every_bench_batch():
    """Sweep of the planning dry run over every tensor-core conv of the full-config path (generator + dense-motion
    hourglass, all four operand formats) at the frame counts the BASELINE configs and the sharded job produce (1 ... 512
    per GPU, odd and ragged ones included): the planner must accept each call and return a launchable plan."""
    code = r"""
import ctypes, json, os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
os.environ["EAMM_TC_NUM_SMS"] = "148"
for k in list(os.environ):
    if k.startswith("EAMM_TC_") and k != "EAMM_TC_NUM_SMS":
        del os.environ[k]
from eamm_b200 import _lib as L
from test_host_logic import _plan
lib = L.load()
C3, UP, C7 = L.CONV_3X3, L.CONV_UP2_3X3, L.CONV_7X7
layers = [("down0", C3, 64, 128, 256, dict(flags=3)), ("down1", C3, 128, 256, 128, dict(flags=3)),
          ("res", C3, 256, 256, 64, dict(flags=1)), ("up0", UP, 256, 128, 64, dict(flags=1)),
          ("up1", UP, 128, 64, 128, dict(flags=1)), ("final", C7, 64, 16, 256, dict(flags=4, nchw_c=3)),
          ("enc0", C3, 64, 128, 64, dict(flags=3)), ("enc1", C3, 128, 256, 32, dict(flags=3)),
          ("enc2", C3, 256, 512, 16, dict(flags=3)), ("enc3", C3, 512, 1024, 8, dict(flags=3)),
          ("enc4", C3, 1024, 1024, 4, dict(flags=3)), ("dec0", UP, 1024, 1024, 2, dict(flags=1)),
          ("dec1", UP, 2048, 512, 4, dict(flags=1)), ("dec2", UP, 1024, 256, 8, dict(flags=1)),
          ("dec3", UP, 512, 128, 16, dict(flags=1)), ("dec4", UP, 256, 64, 32, dict(flags=1)),
          ("mask", C7, 128, 16, 64, dict(nhwc=True))]
out = []
for B in (1, 2, 3, 5, 8, 16, 31, 32, 33, 64, 128, 256, 512):
    for name, kind, cin, cout, hw, kw in layers:
        for planes, f16 in ((2, False), (1, False), (1, True), (2, True)):
            if kind == C7 and f16 and planes == 2:
                continue                       # the 7x7 layers never take mixed operands (engine.py)
            q = _plan(lib, L, kind, cin, cout, B, hw, hw, planes, f16=f16, **kw)
            out.append([name, B, planes, int(f16), q])
# `first` (generator.py:25) as engine.FirstConvTC launches it: packed 16-byte pixels, 7 K chunks, hi/lo or one plane
first = []
for B in (1, 2, 3, 5, 8, 16, 31, 32, 33, 64, 128, 256, 512):
    for H in (64, 256, 512):
        for passes, f16, mix_out in ((2, False, False), (2, False, True), (1, False, False), (1, True, False)):
            inp = L.Act(data=4096, dtype=L.EAMM_F16 if f16 else L.EAMM_BF16, n=B, h=H, w=H, c=8, c_off=0, c_buf=8, planes=1,
                        n_stride=(H + 6) * (H + 8) * 8)
            pl = 2 if passes == 2 else 1
            o = L.Act(data=4096, dtype=L.EAMM_F16 if (f16 or mix_out) else L.EAMM_BF16, n=B, h=H, w=H, c=64, c_off=0,
                      c_buf=64, planes=pl, n_stride=H * H * 64 * pl)
            a = L.ConvArgs()
            a.kind, a.flags, a.cin, a.cout = L.CONV_ROW7_PACKED, L.EPI_RELU, 8, 64
            a.inp, a.out, a.bias, a.weight, a.pack_passes = ctypes.pointer(inp), ctypes.pointer(o), 4096, 4096, passes
            q = (ctypes.c_int * 6)()
            first.append([B, H, passes, lib.eamm_conv_tc_query(ctypes.byref(a), q), list(q)])
print(json.dumps({"layers": out, "first": first}))
""" % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    res = json.loads(r.stdout.strip().splitlines()[-1])
    plans = res["layers"]
    assert len(plans) == 13 * (17 * 4 - 2)
    assert len(res["first"]) == 13 * 3 * 4
    for B, H, passes, rc, q in res["first"]:
        assert rc == 0 and 16 <= q[0] <= 64 and q[5] >= 2, (B, H, passes, rc, q)
        assert q[2] == (2 if passes == 2 else 0), (B, H, passes, q)       # hi/lo planes: fold scheme 2 (include/eamm_b200.h)
    for name, B, planes, f16, q in plans:
        tag = (name, B, planes, f16, q)
        assert 16 <= q["bn"] <= 256 and q["bn"] % 16 == 0, tag
        assert q["stages"] >= 2 and q["ksub"] >= 1, tag
        assert 1 <= q["splitk"] <= 9, tag
        if q["halo_tile"]:                      # fp16 / mixed operands only, never split, never folded
            assert f16 and name not in ("final", "mask") and q["splitk"] == 1 and q["fold"] == 0, tag
        if q["splitk"] > 1:                     # split-K is for the maps whose tiles cannot fill the chip
            assert name.startswith(("enc", "dec", "res", "up", "down")) and not q["pair"], tag
        if name in ("final", "mask"):
            assert q["bn"] == 112 and q["scheme"] in (3, 4), tag
    # the plan is a function of (layer, frames in the launch): equal inputs, equal plans (bit-reproducible results)
    seen = {}
    for name, B, planes, f16, q in plans:
        assert seen.setdefault((name, B, planes, f16), q) == q


def _decode_mix_act(buf):
    """ActBuf in "mix" mode -> (hi, lo8, hi8) fp32 NHWC tensors in the stored (pre-scaled) domain, straight from the bytes."""
    c = buf.c_buf
    hi = buf.t[..., :c].float()
    p1 = buf.t[..., c:].contiguous().view(torch.uint8)
    lo8 = p1[..., :c].view(torch.float8_e4m3fn).float()
    hi8 = p1[..., c:].view(torch.float8_e4m3fn).float()
    return hi, lo8, hi8


def test_mixed_fp16_fp8_operand_format_reproduces_the_conv_on_cpu():
    """The fp16 + 2 x e4m3 operand format (include/eamm_b200.h; conv_tc.cu `mix`): decode the packed weight bytes and an
    encoded activation buffer exactly as the kernel's K loop addresses them (e4m3 lo8 [tap][cin] | e4m3 hi8 [tap][cin] |
    fp16 hi [tap][cin]; a_hi8 x w_lo8 + a_lo8 x w_hi8 + a_hi x w_hi in one accumulator, times acc_scale) and compare with
    the fp32 convolution.  Pins the byte layout, the power-of-two bookkeeping and the accuracy claim (~1e-5 of the output
    scale, against 2.5e-4 for the fp16 term alone)."""
    from eamm_b200 import engine
    g = torch.Generator().manual_seed(3)
    cin, cout, N, H, W = 128, 32, 1, 8, 8
    w = (torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) * (3.0 / (cin * 9)) ** 0.5
    w = w * torch.exp(torch.randn(cout, 1, 1, 1, generator=g))               # rows of very different magnitude
    x = torch.randn(N, H, W, cin, generator=g).relu() * torch.exp(0.8 * torch.randn(N, H, W, cin, generator=g))
    full = w.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous()
    packed, w_exp = engine.pack_tc_weights_mix(full, 1)
    assert packed.shape == (cout, 9 * cin * 4) and packed.dtype == torch.uint8
    assert int((full.abs().amax(dim=(0, 2)) * torch.exp2(w_exp.float())).log2().floor().max()) == 13
    n8 = 9 * cin
    w_lo8 = packed[:, :n8].contiguous().view(torch.float8_e4m3fn).float().view(cout, 9, cin)
    w_hi8 = packed[:, n8:2 * n8].contiguous().view(torch.float8_e4m3fn).float().view(cout, 9, cin)
    w_hi = packed[:, 2 * n8:].contiguous().view(torch.float16).float().view(cout, 9, cin)
    buf = engine.ActBuf(N, H, W, cin, "mix", torch.device("cpu"))
    in_exp = int(engine.ACT_TOP - float(x.abs().max().log2()))
    buf.store_float(x, exp=in_exp)
    assert (buf.to_float().permute(0, 2, 3, 1) - x).abs().max() <= 2.0 ** -15 * x.abs().max()     # ~16 significant bits
    a_hi, a_lo8, a_hi8 = _decode_mix_act(buf)
    assert a_hi.abs().max() < 2.0 ** 12 and a_lo8.abs().max() <= 448 and a_hi8.abs().max() <= 448

    def conv(a, wt):                                   # a NHWC, wt [cout][tap][cin] -> [N,cout,H,W] in fp64
        wk = wt.view(cout, 3, 3, cin).permute(0, 3, 1, 2).double()
        return F.conv2d(a.permute(0, 3, 1, 2).double(), wk, padding=1)

    acc_scale = torch.exp2(-(w_exp.float() + in_exp)).view(1, cout, 1, 1).double()
    main = conv(a_hi, w_hi) * acc_scale
    cross = (conv(a_hi8, w_lo8) + conv(a_lo8, w_hi8)) * acc_scale
    want = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=1)
    scale = want.abs().amax(dim=(0, 2, 3), keepdim=True)                    # per output channel (rows differ by e^3)
    err_main = ((main - want).abs() / scale).max().item()
    err_full = ((main + cross - want).abs() / scale).max().item()
    assert err_full <= 3e-5 and err_main >= 5 * err_full, (err_main, err_full)
    # cin == 64: one fp8 chunk per tap holds both cross terms -- weight bytes [w_hi8 x 64 | w_lo8 x 64] meet the pixel's
    # plane 1 = [a_lo8 x 64 | a_hi8 x 64]
    full64 = full[:, :, :64].contiguous()
    p64, e64 = engine.pack_tc_weights_mix(full64, 1)
    assert p64.shape == (cout, 9 * 64 * 4)
    x8 = p64[:, :9 * 128].contiguous().view(torch.float8_e4m3fn).float().view(cout, 9, 128)
    f16 = p64[:, 9 * 128:].contiguous().view(torch.float16).float().view(cout, 9, 64)
    buf64 = engine.ActBuf(N, H, W, 64, "mix", torch.device("cpu"))
    buf64.store_float(x[..., :64].contiguous(), exp=in_exp)
    a_hi, a_lo8, a_hi8 = _decode_mix_act(buf64)
    plane1 = buf64.t[..., 64:].contiguous().view(torch.uint8).view(torch.float8_e4m3fn).float()       # [.., 128] as the TMA box sees it
    assert torch.equal(plane1[..., :64], a_lo8) and torch.equal(plane1[..., 64:], a_hi8)

    def conv_k(a, wt, k):
        wk = wt.view(cout, 3, 3, k).permute(0, 3, 1, 2).double()
        return F.conv2d(a.permute(0, 3, 1, 2).double(), wk, padding=1)

    sc64 = torch.exp2(-(e64.float() + in_exp)).view(1, cout, 1, 1).double()
    got64 = (conv_k(plane1, x8, 128) + conv_k(a_hi, f16, 64)) * sc64
    want64 = F.conv2d(x[..., :64].permute(0, 3, 1, 2).double(), w[:, :64].double(), padding=1)
    s64 = want64.abs().amax(dim=(0, 2, 3), keepdim=True)
    assert ((got64 - want64).abs() / s64).max().item() <= 3e-5


def test_numpy_sampler_matches_torch_grid_sample_and_corner_values():
    assert sampler_np.unnormalize(np.float32(-1.0), 64) == -0.5          # SURVEY.md §7 "hard parts"
    assert sampler_np.unnormalize(np.float32(1.0), 64) == 63.5
    g = torch.Generator().manual_seed(3)
    img = torch.rand(3, 8, 8, generator=g)
    grid = torch.rand(8, 8, 2, generator=g) * 2.6 - 1.3                     # includes out-of-range samples
    want = F.grid_sample(img[None], grid[None], align_corners=False)[0].numpy()
    got = sampler_np.grid_sample(img.numpy(), grid.numpy())
    np.testing.assert_allclose(got, want, atol=1e-6)
    ident = oracle.make_coordinate_grid(8, 8).numpy()                       # the "identity" grid is not a copy
    y0, x0, w, ok = sampler_np.grid_sample_taps(ident[0, 0, 0], ident[0, 0, 1], 8, 8)[0]
    assert (y0, x0, ok) == (-1, -1, False)


def test_bilinear_and_nearest_index_rules():
    flow = torch.arange(16, dtype=torch.float32).view(1, 1, 4, 4)
    up = F.interpolate(flow, size=(16, 16), mode="bilinear", align_corners=False)[0, 0]
    for dst in (0, 1, 2, 7, 14, 15):
        i0, i1, lam = sampler_np.bilinear_upsample_index(dst, 4, 16)
        want = flow[0, 0, 0, i0] * (1 - lam) + flow[0, 0, 0, i1] * lam
        assert abs(up[0, dst].item() - want.item()) < 1e-6
    near = F.interpolate(flow, scale_factor=2)[0, 0]
    for dst in range(8):
        assert near[0, dst] == flow[0, 0, 0, sampler_np.nearest_up2_index(dst)]


def test_channel_interleave_and_emotion_row_selection_are_exact():
    """hourglass input channel order [hm_k, R_k, G_k, B_k] (dense_motion.py:93-94); demo.py:266-271 rows."""
    cfg = get_config("tiny")
    sd = synth.make_state_dict(cfg)
    src, kpd, kps = synth.make_inputs(1, cfg, size=64, seed=5)
    taps = {}
    oracle.generator_forward(sd, cfg, src, kpd, kps, taps=taps)
    hin, hm = taps["hourglass_in"], taps["heatmap"]
    K1 = cfg["num_kp"] + 1
    for k in range(K1):
        assert torch.equal(hin[:, 4 * k], hm[:, k, 0])
    value = torch.zeros(1, 10, 2)
    emo = torch.arange(8, dtype=torch.float32).view(1, 4, 2) + 1
    for dst, srcrow, scale in ((1, 0, 1.0), (4, 1, 0.2), (6, 2, 1.0)):     # demo.py:266-271
        value[:, dst] += emo[:, srcrow] * scale
    assert value[0, 1].tolist() == [1.0, 2.0] and value[0, 6].tolist() == [5.0, 6.0]
    assert torch.allclose(value[0, 4], torch.tensor([0.6, 0.8]))


# ------------------------------------------------------------------ drop-in module surface
def test_dropin_modules_keep_reference_state_dict_layout_and_refuse_cpu():
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    from eamm_b200.modules.dense_motion import DenseMotionNetwork
    cfg = get_config("full")
    gen = OcclusionAwareGenerator(**cfg).eval()
    sd = synth.make_state_dict(cfg)
    assert len(sd) == 196
    res = gen.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert list(gen.state_dict().keys())[:2] == ["dense_motion_network.hourglass.encoder.down_blocks.0.conv.weight",
                                                 "dense_motion_network.hourglass.encoder.down_blocks.0.conv.bias"]
    assert gen.num_channels == 3 and isinstance(gen.dense_motion_network, DenseMotionNetwork)
    assert "OcclusionAwareGenerator" in repr(gen)
    src, kpd, kps = synth.make_inputs(1, cfg, size=256)
    with pytest.raises(RuntimeError, match="no CPU path"):
        gen(src, kp_driving=kpd, kp_source=kps)
    gen.train()
    with pytest.raises(RuntimeError, match="inference path only"):
        gen(src, kp_driving=kpd, kp_source=kps)
    with pytest.raises(ValueError):
        gen.precision = "fp64"
    # the packed-weight engine is dropped (parent and dense-motion child) by everything that may change the parameters
    # (the child's own engine only exists when DenseMotionNetwork is called stand-alone; it has its own `precision`)
    for change, child_too in ((lambda: gen.load_state_dict(sd), True), (lambda: gen.float(), True),
                              (lambda: gen.refresh_weights(), True), (lambda: setattr(gen, "precision", "fp16"), False)):
        gen._eng = gen.dense_motion_network._eng = object()
        change()
        assert gen._eng is None
        assert (gen.dense_motion_network._eng is None) == child_too


def test_kp_detector_and_at_net2_dropins_keep_the_reference_parameter_layout():
    """the drop-ins load, strictly, the seeded state dicts that tools/make_golden.py loaded
    strictly into the REAL KPDetector / KPDetector_a / AT_net2 (that is what pins the key names and shapes); AT_net2 also
    swallows the unused StyleGAN2 `generator.*` entries of a reference checkpoint; all of them refuse to run on the CPU."""
    from eamm_b200.config import get_kp_config
    from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
    from eamm_b200.modules.util import AT_net2
    for audio, cls in ((False, KPDetector), (True, KPDetector_a)):
        cfg = get_kp_config("full", audio=audio)
        det = cls(**cfg).eval()
        sd = synth.make_kp_state_dict(cfg)
        res = det.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        assert set(det.state_dict()) == set(sd)
        x = synth.make_kp_inputs(cfg, 1, 256, audio)
        with pytest.raises(RuntimeError, match="no CPU path"):
            det(x)
    at = AT_net2().eval()
    sd = synth.make_at_state_dict()
    sd["generator.style.1.weight"] = torch.zeros(4, 4)               # a reference checkpoint carries the StyleGAN2 decoder
    res = at.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert not any(k.startswith("generator.") for k in at.state_dict())
    img, mfcc, pose = synth.make_at_inputs(1, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        at(img, mfcc, pose, "cnn", 1.6)
    with pytest.raises(NotImplementedError):
        at(img, mfcc, pose, "gan", 1.6)
    with pytest.raises(Exception, match="jaco_net type wrong"):       # util.py:611
        at(img, mfcc, pose, "other", 1.6)


def test_conv_layer_table_matches_reference_flop_count():
    cfg = get_config("full")
    flops = 0.0
    res = {"dense": 64, "first": 256, "final": 256}
    for prefix, cin, cout, k, kind in synth.conv_layers(cfg):
        if "hourglass.encoder" in prefix:
            hw = 64 >> int(prefix[-1])
        elif "hourglass.decoder" in prefix:
            hw = 4 << int(prefix[-1])
        elif prefix.startswith("dense_motion"):
            hw = 64
        elif prefix.startswith("down_blocks"):
            hw = 256 >> int(prefix[-1])
        elif prefix.startswith("up_blocks"):
            hw = 128 << int(prefix[-1])
        elif prefix.startswith("bottleneck"):
            hw = 64
        else:
            hw = 256
        n = 2 if kind == "res" else 1
        flops += n * 2.0 * cin * cout * k * k * hw * hw
    assert abs(flops / 1e9 - 107.286) < 0.01


# ------------------------------------------------------------------ C ABI
def test_shared_library_exports_every_declared_symbol():
    from eamm_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "eamm_b200.h")).read()
    declared = set(re.findall(r"\bint\s+(eamm_\w+)\s*\(", header))
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name
    lib.eamm_abi_version.restype = ctypes.c_int
    assert lib.eamm_abi_version() == 2
    null = ctypes.c_void_p(0)
    lib.eamm_warp_image.restype = ctypes.c_int
    assert lib.eamm_warp_image(null, 0, null, null, 1, 3, 8, 8, 2, 2, null) == -1      # EAMM_ERR_ARG, no launch


def test_every_entry_point_rejects_null_arguments_without_touching_a_device():
    """Error convention of the boundary (include/eamm_b200.h: < 0 = EAMM_ERR_*, rejected before anything is launched): every
    launching entry point returns EAMM_ERR_ARG on null buffers / zeroed structs, on a host with or without a GPU; the pure
    planning helpers answer from the layer shape alone."""
    from eamm_b200 import _lib as L
    lib = L.load()
    C = ctypes
    act, kp, ca = L.Act(), L.Kp(), L.ConvArgs()
    n = None
    calls = {
        "eamm_aa_downsample": lambda: lib.eamm_aa_downsample(n, 0, n, 1, 256, 256, 4, n, 13, n),
        "eamm_aa_downsample_act": lambda: lib.eamm_aa_downsample_act(n, 0, 1, 256, 256, 4, n, 13, C.byref(act), n),
        "eamm_kp_head": lambda: lib.eamm_kp_head(n, 16, 1, 64, 64, 10, 10, 3, 0.1, n, n, n, n),
        "eamm_kp_clip": lambda: lib.eamm_kp_clip(n, n, n, n, 4, 10, 0, n, n, n, n, 0, n, n, n, n, 1.0, 1, n, n, n, n),
        "eamm_kp_stage": lambda: lib.eamm_kp_stage(n, 0, C.byref(kp), C.byref(kp), 10, 0.01, C.byref(act), n, n, n),
        "eamm_flow_combine": lambda: lib.eamm_flow_combine(n, 16, C.byref(kp), C.byref(kp), 10, 1, 1, 64, 64, n, n, n, n),
        "eamm_warp_occlude": lambda: lib.eamm_warp_occlude(C.byref(act), n, n, 0, 0, C.byref(act), None, n, n, n, n),
        "eamm_warp_image": lambda: lib.eamm_warp_image(n, 0, n, n, 1, 3, 8, 8, 2, 2, n),
        "eamm_nchw_to_act": lambda: lib.eamm_nchw_to_act(n, 1, 3, 8, 8, C.byref(act), n),
        "eamm_conv_simt": lambda: lib.eamm_conv_simt(C.byref(ca), n),
        "eamm_conv_tc": lambda: lib.eamm_conv_tc(C.byref(ca), n),
        "eamm_conv_tc_query": lambda: lib.eamm_conv_tc_query(C.byref(ca), (C.c_int * 6)()),
        "eamm_pack_image": lambda: lib.eamm_pack_image(n, 1, 3, 8, 8, 0, n, n),
        "eamm_linear": lambda: lib.eamm_linear(n, 4, n, n, n, 1, n, 4, 1, 4, 4, 0, 1.0, n),
        "eamm_maxpool": lambda: lib.eamm_maxpool(C.byref(act), C.byref(act), 3, 1, 2, n),
        "eamm_act_copy": lambda: lib.eamm_act_copy(C.byref(act), C.byref(act), 8, 8, 0, n),
        "eamm_lstm_layer": lambda: lib.eamm_lstm_layer(n, n, n, 1, 1, 256, n),
    }
    helpers = {"eamm_abi_version", "eamm_device_ok", "eamm_conv_tc_uses_halo", "eamm_conv_tc_fold"}
    assert set(calls) | helpers == set(L.exported_symbols())          # a new entry point must be added here
    before = L.LAUNCHES
    for name, fn in calls.items():
        assert fn() == -1, name                                         # EAMM_ERR_ARG
        with pytest.raises(RuntimeError, match="EAMM_ERR_ARG"):
            L.check(fn(), name)
    assert L.LAUNCHES == before                                         # rejected calls are not counted as launches
    # weight-packing helpers (include/eamm_b200.h): hi/lo layers with cout <= 128 stack the weight planes along N (fold 1),
    # cout = 256 keeps K = (pass, tap, channel); the packed first conv folds as scheme 2; one-plane layers never fold
    assert [lib.eamm_conv_tc_fold(L.CONV_3X3, 1, c, 0) for c in (64, 128, 256)] == [1, 1, 0]
    assert [lib.eamm_conv_tc_fold(L.CONV_3X3, 0, c, 0) for c in (64, 128, 256)] == [0, 0, 0]
    assert (lib.eamm_conv_tc_fold(L.CONV_ROW7_PACKED, 1, 64, 0), lib.eamm_conv_tc_fold(L.CONV_ROW7_PACKED, 0, 64, 0)) == (2, 0)
    assert lib.eamm_conv_tc_uses_halo(L.CONV_3X3, 64, 256, 0) == 0      # 7x7 schemes only
    assert lib.eamm_conv_tc_uses_halo(L.CONV_7X7, 256, 16, 3) == 2      # `final`: kx in N (the query may widen it to scheme 3)


def test_conv_tc_argument_validation_codes():
    """eamm_conv_tc validates before it plans (the query runs the same checks and launches nothing): each class of bad call
    maps to its EAMM_ERR_* code, so a binding can tell a caller error from a device error (> 0 = cudaError_t)."""
    code = r"""
import ctypes as C, json, os, sys
sys.path.insert(0, %r)
os.environ["EAMM_TC_NUM_SMS"] = "148"
from eamm_b200 import _lib as L
lib = L.load()
def q(kind=L.CONV_3X3, cin=64, cout=64, n=2, h=16, w=16, planes=1, dt=L.EAMM_BF16, data=4096, wptr=4096, oh=None, c_buf=None):
    a = L.ConvArgs(); a.kind, a.flags, a.cin, a.cout = kind, 0, cin, cout
    cb = c_buf or cin
    act = L.Act(data=data, dtype=dt, n=n, h=h, w=w, c=cin, c_off=0, c_buf=cb, planes=planes, n_stride=h * w * planes * cb)
    o = oh or h
    out = L.Act(data=4096, dtype=dt, n=n, h=o, w=o, c=cout, c_off=0, c_buf=cout, planes=planes, n_stride=o * o * planes * cout)
    a.inp, a.weight, a.bias, a.out = C.pointer(act), wptr, 4096, C.pointer(out)
    a.splitk_ws, a.splitk_ws_bytes = 4096, L.SPLITK_WS_BYTES
    return lib.eamm_conv_tc_query(C.byref(a), (C.c_int * 6)())
print(json.dumps({"ok": q(), "map6": q(h=6, w=6), "kind": q(kind=9), "cin32": q(cin=32), "cin96": q(cin=96, c_buf=128),
                  "act_align": q(data=4100), "w_align": q(wptr=4100), "f32": q(dt=L.EAMM_F32), "planes3": q(planes=3),
                  "out_size": q(oh=8), "cout20": q(cout=20), "n0": q(n=0)}))
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    rc = json.loads(r.stdout.strip().splitlines()[-1])
    ARG, SHAPE, DTYPE, ALIGN, UNSUPPORTED = -1, -2, -3, -4, -5
    assert rc["ok"] == 0
    assert rc["map6"] == UNSUPPORTED and rc["kind"] == UNSUPPORTED      # non power-of-two maps, unknown conv kind
    assert rc["cin32"] == ALIGN and rc["cin96"] == ALIGN                 # K chunks are 64 channels = one 128-byte swizzled row
    assert rc["act_align"] == ALIGN and rc["w_align"] == ALIGN           # TMA needs 16-byte aligned bases
    assert rc["f32"] == DTYPE and rc["planes3"] == DTYPE                 # tensor-core operands: bf16 / fp16, one or two planes
    assert rc["out_size"] == SHAPE and rc["cout20"] == SHAPE             # output view must match the layer
    assert rc["n0"] == ARG


def test_epilogue_chunk_walk_covers_every_column_chunk_once():
    """Epilogue of conv_tc.cu (epilogue_tile / the fast variants): the two epilogue warps of a TMEM lane quadrant start at
    chunk `half` and step by two chunks; every accumulator column chunk is read exactly once, for every N tile the planner
    produces (the split-K reduction by the last-arriving CTA walks the workspace the same way)."""
    halves = 2
    for bn, ch in ((256, 32), (128, 32), (112, 16), (64, 32), (48, 16), (32, 32), (16, 16)):
        seen = [c0 // ch for half in range(halves) for c0 in range(half * ch, bn, halves * ch)]     # cfirst, cstep
        assert sorted(seen) == list(range(bn // ch)), (bn, ch, seen)


def test_ctypes_structs_mirror_the_header_layout(tmp_path):
    """The ctypes mirrors in eamm_b200/_lib.py have the field offsets and sizes gcc gives include/eamm_b200.h."""
    from eamm_b200 import _lib as L
    structs = {"eamm_act": L.Act, "eamm_kp": L.Kp, "eamm_conv_args": L.ConvArgs, "eamm_one_euro": L.OneEuro}
    rename = {"inp": "in"}                       # `in` is a Python keyword
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "eamm_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, rename.get(fname, fname)))
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, cls in structs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    assert L.SPLITK_WS_BYTES == 4096 + 160 * 128 * 256 * 4
    hdr = open(os.path.join(ROOT, "include", "eamm_b200.h")).read()
    assert "#define EAMM_SPLITK_WS_BYTES (4096 + 160ll * 128 * 256 * 4)" in hdr


# ------------------------------------------------------------------ sharding (gloo, world_size 2)
def test_partition_covers_every_frame_once():
    for total in (0, 1, 7, 32, 1024):
        for world in (1, 2, 3, 8):
            spans = [sharding.partition(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from eamm_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD"]))
rank, world = dist.get_rank(), dist.get_world_size()
total = int(os.environ["TOTAL"])
a, b = sharding.partition(total, world, rank)
frames = torch.arange(a, b, dtype=torch.float32).view(-1, 1) * 10.0       # "generated frames" of this rank
allf = sharding.gather_frames(frames, total, dst=0)
mx = sharding.reduce_max(1.0 + rank)
sm = sharding.reduce_sum(float(b - a))
if rank == 0:
    assert allf.view(-1).tolist() == [10.0 * i for i in range(total)], allf
    print("OK", mx, sm)
assert mx == float(world) and sm == total
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world,total", [(2, 7), (4, 1026), (3, 2)])     # uneven blocks; a rank that owns no frame at all
def test_gloo_multi_rank_gather_and_max_reduce(tmp_path, world, total):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), PORT=str(port), WORLD=str(world), TOTAL=str(total))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK %.1f %.1f" % (world, total) in outs[0]


# ------------------------------------------------------------------ AT_net2 packing
def _emulate_up2(layer, x):
    """What the UP2 kernels compute from ConvLayer.w_ref ([16][cout][cin]): out(2y+a, 2x+b) = sum over taps."""
    cout = layer.cout
    pw = layer.w_ref.view(4, 4, cout, -1)
    N, _, H, W = x.shape
    out = torch.zeros(N, cout, 2 * H, 2 * W)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in (0, 1):
        for b in (0, 1):
            acc = layer.bias.view(1, -1, 1, 1)
            for ty in (0, 1):
                for tx in (0, 1):
                    dy, dx = a - 1 + ty, b - 1 + tx
                    acc = acc + torch.einsum("oc,nchw->nohw", pw[a * 2 + b, ty * 2 + tx], xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W])
            out[:, :, a::2, b::2] = acc
    return F.relu(out) if layer.flags & 1 else out


def _emulate_conv3(layer, x):
    w = layer.w_ref.view(3, 3, layer.cout, -1).permute(2, 3, 0, 1)
    y = F.conv2d(F.pad(x, (0, 0, 0, 0, 0, w.shape[1] - x.shape[1])), w, layer.bias, padding=1)
    y = F.relu(y) if layer.flags & 1 else y
    return F.avg_pool2d(y, 2) if layer.flags & 2 else y


def _emulate_conv7(layer, x):
    w = layer.w_ref.view(7, 7, layer.cout, -1).permute(2, 3, 0, 1)
    y = F.conv2d(F.pad(x, (0, 0, 0, 0, 0, w.shape[1] - x.shape[1])), w, layer.bias, padding=3)
    y = F.relu(y) if layer.flags & 1 else y
    return torch.sigmoid(y) if layer.flags & 4 else y


def _emulate_hourglass(hg, x):
    """HourglassPlan.run with torch: level-L buffer = [up-block slot | skip slot]; producers write their slot, consumers read
    a slot (encoder) or the whole buffer (decoder, head).  Returns cat[0] = [decoder output | input] (padded slots)."""
    nb, ca = hg.nb, hg.calign
    ru = lambda c: (c + ca - 1) // ca * ca
    B, _, h, w = x.shape
    s_up = [ru(hg.dec_ch[nb - 1 - l]) for l in range(nb)]
    s_sk = [ru(hg.enc_ch[l]) for l in range(nb)]
    cat = [torch.zeros(B, s_up[l] + s_sk[l], h >> l, w >> l) for l in range(nb)]
    bott = torch.zeros(B, ru(hg.enc_ch[nb]), h >> nb, w >> nb)
    cat[0][:, s_up[0]:s_up[0] + x.shape[1]] = x
    for i, layer in enumerate(hg.enc_layers):
        assert layer.cin == s_sk[i]
        y = _emulate_conv3(layer, cat[i][:, s_up[i]:s_up[i] + s_sk[i]])
        if i + 1 < nb:
            cat[i + 1][:, s_up[i + 1]:s_up[i + 1] + layer.cout] = y
        else:
            bott[:, :layer.cout] = y
    for j, layer in enumerate(hg.dec_layers):
        src = bott if j == 0 else cat[nb - j]
        assert layer.cin == src.shape[1]
        cat[nb - 1 - j][:, :layer.cout] = _emulate_up2(layer, src)
    return cat[0], s_up[0]


@pytest.mark.parametrize("cfg_name", ["tiny", "tiny_sf1"])
def test_generator_engine_packing_reproduces_the_oracle_on_cpu(cfg_name):
    """Re-executes the PACKED generator (engine.GeneratorEngine / DenseMotionEngine / HourglassPlan built on the CPU: folded
    BatchNorm, pre-activation affines moved into the producers' epilogues, merged mask + occlusion conv, UP2 parity weights,
    channel-slot buffers instead of torch.cat, the 1-D anti-alias factor) with plain torch ops in the order engine._forward
    launches the kernels, and compares with the oracle: every host-side rewrite of generator.py:59-97 / dense_motion.py:81-113
    is pinned without a GPU.  The kernels' own arithmetic (heatmaps, sampling, softmax) is taken from the oracle here."""
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg = get_config(cfg_name)
    sd = synth.make_state_dict(cfg, seed=0)
    m = OcclusionAwareGenerator(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        eng = engine.GeneratorEngine(m, "fp32_simt")
        assert not eng.mixed and eng.dm is not None and not eng.first_packed
        B, size = 2, 64
        src, kpd, kps = synth.make_inputs(B, cfg, size=size, seed=1)
        want = oracle.generator_forward(sd, cfg, src, kpd, kps)
        # encoder
        x = _emulate_conv7(eng.first, src)
        for layer in eng.down:
            x = _emulate_conv3(layer, x)
        feat = x
        # dense motion
        dm = eng.dm
        small = src
        if dm.step != 1:
            k2 = torch.outer(dm.g1, dm.g1).expand(3, 1, dm.taps, dm.taps)
            pad = dm.taps // 2
            small = F.conv2d(F.pad(src, (pad,) * 4), k2, groups=3)[:, :, ::dm.step, ::dm.step]
        h, w = small.shape[2:]
        dmp = cfg["dense_motion_params"]
        hm = oracle.heatmap_representation(kpd, kps, h, w, dmp.get("kp_variance", 0.01))
        sm = oracle.sparse_motions(kpd, kps, h, w)
        ds = oracle.deformed_source(small, sm)
        hg_in = torch.cat([hm, ds], dim=2).view(B, -1, h, w)                     # [hm_k, R_k, G_k, B_k] per keypoint
        cat0, s_up0 = _emulate_hourglass(dm.hg, hg_in)
        assert dm.head.cin == cat0.shape[1]
        logits = _emulate_conv7(dm.head, cat0)                                     # merged mask (K+1) + occlusion (1) conv
        K1 = cfg["num_kp"] + 1
        mask = F.softmax(logits[:, :K1], dim=1)
        occ = torch.sigmoid(logits[:, K1:K1 + 1])
        deformation = (sm.permute(0, 1, 4, 2, 3) * mask.unsqueeze(2)).sum(dim=1).permute(0, 2, 3, 1)
        assert (mask - want["mask"]).abs().max() <= 2e-5 and (occ - want["occlusion_map"]).abs().max() <= 2e-5
        # warp x occlusion, then the bottleneck with every BatchNorm living in a neighbouring epilogue
        x = oracle.deform_input(feat, deformation)
        o = occ if occ.shape[2:] == x.shape[2:] else F.interpolate(occ, size=x.shape[2:], mode="bilinear", align_corners=False)
        x = x * o
        c = x.shape[1]
        a = F.relu(x * eng.pre[0][:c].view(1, -1, 1, 1) + eng.pre[1][:c].view(1, -1, 1, 1))      # norm1 / ReLU of block 0
        for i, (l1, l2) in enumerate(eng.res):
            t = _emulate_conv3(l1, a)                                              # conv1 + folded norm2 + ReLU
            x = _emulate_conv3(l2, t) + x                                          # conv2 + residual
            if l2.scale2 is not None:                                              # next block's norm1 / ReLU (second output)
                a = F.relu(x * l2.scale2.view(1, -1, 1, 1) + l2.shift2.view(1, -1, 1, 1))
        assert eng.res[-1][1].scale2 is None
        for layer in eng.up:
            x = _emulate_up2(layer, x)
        pred = _emulate_conv7(eng.final, x)[:, :3]
        err = (pred - want["prediction"]).abs().max().item()
        assert err <= 2e-5, err


@pytest.mark.parametrize("audio", [False, True])
def test_kp_detector_engine_packing_reproduces_the_oracle_on_cpu(audio):
    """SURVEY 8(f) rank 1, host side: KPDetectorEngine built on the CPU (Hourglass predictor in slot buffers, keypoint and
    Jacobian 7x7 convs merged into ONE same-padded conv whose interior is the reference's valid conv,
    keypoint_detector.py:82-103 / :183-203) re-executed with torch ops equals the oracle's value / jacobian / heatmap."""
    from eamm_b200.config import get_kp_config
    from eamm_b200.kp_engine import KPDetectorEngine
    from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg = get_kp_config("tiny", audio=audio)
    sd = synth.make_kp_state_dict(cfg)
    m = (KPDetector_a if audio else KPDetector)(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    B, size = 2, 64
    x = synth.make_kp_inputs(cfg, B, size, audio)
    with torch.no_grad():
        eng = KPDetectorEngine(m, "fp32_simt")
        want = (oracle.kp_detector_a_forward if audio else oracle.kp_detector_forward)(sd, cfg, x)
        if eng.hg is not None:
            small = x
            if eng.step != 1:
                k2 = torch.outer(eng.g1, eng.g1).expand(3, 1, eng.taps, eng.taps)
                small = F.conv2d(F.pad(x, (eng.taps // 2,) * 4), k2, groups=3)[:, :, ::eng.step, ::eng.step]
            feat, _ = _emulate_hourglass(eng.hg, small)
        else:
            feat = x
        logits = _emulate_conv7(eng.head, feat)                 # same-padded: [B, K + 4J (padded), h, w]
        off = 3 - m.pad
        if off:
            logits = logits[:, :, off:-off, off:-off]           # the valid conv of the reference
        K, J = m.num_kp, m.num_jacobian_maps
        shape = (B, K) + tuple(logits.shape[2:])
        heat = F.softmax(logits[:, :K].reshape(B, K, -1) / m.temperature, dim=2).view(*shape)
        assert heat.shape == want["heatmap"].shape and (heat - want["heatmap"]).abs().max() <= 1e-5
        assert (oracle.gaussian2kp(heat) - want["value"]).abs().max() <= 1e-5
        if J:
            jm = logits[:, K:K + 4 * J].reshape(B, J, 4, *shape[2:])
            jac = (heat.unsqueeze(2) * jm).view(B, K, 4, -1).sum(dim=-1).view(B, K, 2, 2)
            assert (jac - want["jacobian"]).abs().max() <= 1e-4


def test_operand_format_plan_of_the_full_config(monkeypatch):
    """Which layers run which operand scheme in each precision mode, read off engines built on the
    CPU for the BASELINE configuration: fp32 mode = fp16 + 2 x e4m3 on every 3x3 / UP2 conv of the generator (down0/1, the 12
    bottleneck convs, up0/1), bf16 hi/lo on the 7x7 convs and the dense-motion Hourglass; 17 calibrated tensors; the switches
    documented in INTEGRATION.md change exactly what they say."""
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    for k in ("EAMM_B200_MIX", "EAMM_B200_MIX_HG", "EAMM_B200_MIX_SKIP", "EAMM_B200_MIX64", "EAMM_B200_CONV", "EAMM_TC_ROW7"):
        monkeypatch.delenv(k, raising=False)
    cfg = get_config("full")
    m = OcclusionAwareGenerator(**cfg).eval()
    m.load_state_dict(synth.make_state_dict(cfg, seed=0))

    def plan(precision):
        with torch.no_grad():
            e = engine.GeneratorEngine(m, precision)
        gen = [l.impl for l in e.down] + [l.impl for pair in e.res for l in pair] + [l.impl for l in e.up]
        hg = [l.impl for l in e.dm.hg.enc_layers + e.dm.hg.dec_layers]
        return e, gen, hg

    e, gen, hg = plan("fp32")
    assert e.mixed and e.first_packed and gen == ["mix"] * 16 and hg == ["tc3"] * 10
    assert (e.final.impl, e.dm.head.impl, e.first.split, e.first.f16) == ("tc3", "tc3", True, False)
    assert (e.enc_mix, e.res_mix, e.xf_mix, e.dec_mix) == ([True, True, True], True, True, [True, False])
    assert sorted(e.calib.slots) == sorted(["enc0", "enc1", "enc2", "xf", "dec0"] + ["a%d" % i for i in range(6)] +
                                           ["t%d" % i for i in range(6)])
    assert all(l.w_exp.shape == (l.cout,) and l.weight.dtype == torch.uint8 for l in e.down + e.up)
    for precision, impl, mode in (("fp32_bf16x3", "tc3", "bf16x2"), ("fp16", "tc16", "f16"), ("bf16", "tc", "bf16"),
                                  ("fp32_simt", "simt", "f32")):
        e, gen, hg = plan(precision)
        assert not e.mixed and e.calib is None and (e.impl, e.mode) == (impl, mode)
        assert gen == [impl] * 16 and hg == [impl] * 10 and e.final.impl == impl
        assert e.first_packed == (impl != "simt")
    monkeypatch.setenv("EAMM_B200_MIX", "0")                      # identical to fp32_bf16x3
    e, gen, hg = plan("fp32")
    assert not e.mixed and gen == ["tc3"] * 16
    monkeypatch.setenv("EAMM_B200_MIX", "res")                    # the bottleneck only (the first cut of the scheme)
    e, gen, hg = plan("fp32")
    assert gen == ["tc3"] * 2 + ["mix"] * 12 + ["tc3"] * 2 and sorted(e.calib.slots) == sorted(
        ["a%d" % i for i in range(6)] + ["t%d" % i for i in range(6)])
    monkeypatch.setenv("EAMM_B200_MIX", "1")
    monkeypatch.setenv("EAMM_B200_MIX_HG", "1")                   # opt-in: eligible Hourglass layers too (error table in profiles/)
    e, gen, hg = plan("fp32")
    assert gen == ["mix"] * 16 and "mix" in hg and hg[0] == "tc3"   # enc0 reads the 44-channel input: never mixed


def test_every_tensor_core_weight_matrix_of_an_engine_decodes_to_its_fp32_weights():
    """Every eamm_conv_tc weight matrix a GeneratorEngine holds (modes fp32_bf16x3 / fp16 / bf16 and the mixed fp16 + e4m3
    layers of fp32 mode), decoded from its documented layout (include/eamm_b200.h: rows = (class, cout), K = (pass, tap,
    channel) with planes lo, hi, hi; mixed: e4m3 lo8 | e4m3 hi8 | fp16 hi, one exponent per cout), gives back the layer's fp32
    weights to the precision of the format -- for 3x3 and UP2 layers, padded couts / channel slots included."""
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    cfg = get_config("tiny")
    m = OcclusionAwareGenerator(**cfg).eval()
    m.load_state_dict(synth.make_state_dict(cfg, seed=0))

    def layers(e):
        out = list(e.down) + [l for pair in e.res for l in pair] + list(e.up)
        return out + e.dm.hg.enc_layers + e.dm.hg.dec_layers

    for precision, passes, rel in (("fp32_bf16x3", 3, 2.0 ** -15), ("fp16", 1, 2.0 ** -10), ("bf16", 1, 2.0 ** -7)):
        with torch.no_grad():
            e = engine.GeneratorEngine(m, precision)
        seen = 0
        for l in layers(e):
            classes = 4 if l.kind == L_UP2 else 1
            taps = l.w_ref.shape[0] // classes
            ref = l.w_ref.view(classes, taps, l.cout, l.cin).permute(0, 2, 1, 3)                 # [cls][cout][tap][cin]
            W = l.weight.float().view(classes, l.cout, passes, taps, l.cin)
            if passes == 3:
                assert torch.equal(W[:, :, 1], W[:, :, 2])                                         # hi plane twice (a_lo, a_hi passes)
                dec = W[:, :, 0] + W[:, :, 1]
            else:
                dec = W[:, :, 0]
            scale = ref.abs().amax().item()
            assert (dec - ref).abs().max().item() <= rel * scale, (precision, l.name)
            seen += 1
        assert seen == len(e.down) + 2 * len(e.res) + len(e.up) + 2 * e.dm.hg.nb
    # mixed layers of fp32 mode (tiny config: whichever layers are eligible; the full config's are pinned by the plan test)
    cfgf = get_config("full")
    mf = OcclusionAwareGenerator(**cfgf).eval()
    mf.load_state_dict(synth.make_state_dict(cfgf, seed=0))
    with torch.no_grad():
        e = engine.GeneratorEngine(mf, "fp32")
    mixed = [l for l in layers(e) if l.impl == "mix"]
    assert len(mixed) == 16
    for l in (mixed[0], mixed[1], mixed[2], mixed[-2], mixed[-1]):                                 # down0 (cin 64), down1, res, up0, up1
        classes = 4 if l.kind == L_UP2 else 1
        taps = l.w_ref.shape[0] // classes
        rows, kk = classes * l.cout, taps * l.cin
        ref = l.w_ref.view(classes, taps, l.cout, l.cin).permute(0, 2, 1, 3).reshape(rows, taps, l.cin)
        raw = l.weight
        assert raw.shape == (rows, 4 * kk)
        hi = raw[:, 2 * kk:].contiguous().view(torch.float16).float().view(rows, taps, l.cin)
        if l.cin == 64:                                   # per tap one 128-byte fp8 chunk [hi8 x 64 | lo8 x 64]
            x8 = raw[:, :2 * kk].contiguous().view(torch.float8_e4m3fn).float().view(rows, taps, 128)
            hi8, lo8 = x8[..., :64], x8[..., 64:]
        else:
            lo8 = raw[:, :kk].contiguous().view(torch.float8_e4m3fn).float().view(rows, taps, l.cin)
            hi8 = raw[:, kk:2 * kk].contiguous().view(torch.float8_e4m3fn).float().view(rows, taps, l.cin)
        sc = torch.exp2(-l.w_exp.float()).repeat(classes).view(rows, 1, 1)
        dec = (hi + lo8 / 64.0) * sc
        scale = ref.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-30)
        assert ((dec - ref).abs() / scale).max().item() <= 2.0 ** -13, l.name                     # 11 + 4 bits, per-row scale
        assert ((hi8 * 64.0 - hi).abs() / hi.abs().amax(dim=(1, 2), keepdim=True)).max().item() <= 2.0 ** -4   # hi8 = e4m3(hi / 64)


class _DryLib:
    """The C ABI with every launch replaced by its argument check: eamm_conv_tc -> eamm_conv_tc_query (the same validation and
    planning code, nothing launched), every other kernel entry point -> a recorder that returns 0."""

    def __init__(self, real, log):
        self._real, self._log = real, log

    def __getattr__(self, name):
        real, log = self._real, self._log
        if name == "eamm_conv_tc":
            def conv_tc(args, stream):
                q = (ctypes.c_int * 6)()
                rc = real.eamm_conv_tc_query(args, q)
                log.append(("conv_tc", rc, args._obj.inp.contents.n, tuple(q)))
                return rc
            return conv_tc
        if name in ("eamm_conv_tc_query", "eamm_conv_tc_fold", "eamm_conv_tc_uses_halo", "eamm_abi_version"):
            return getattr(real, name)

        def stub(*a):
            log.append((name[5:], 0, None, None))
            return 0
        return stub


FRAME_LAUNCHES = (["pack_image", "conv_tc", "conv_tc", "conv_tc", "aa_downsample", "kp_stage"] + ["conv_tc"] * 11 +
                  ["flow_combine", "warp_occlude", "warp_image"] + ["conv_tc"] * 15)


def test_host_program_dry_run_launch_sequence_and_argument_checks(monkeypatch):
    """engine._forward end to end on the CPU with the launches stubbed out (_DryLib): the host program of
    generator.py:59-97 / dense_motion.py:81-113 issues the 35 launches DESIGN.md lists, in that order, and all 29 eamm_conv_tc
    argument structs it builds -- real buffers, slot views, operand formats, pre-scale exponents, weight packings chosen from the
    planner's answer -- pass the library's own validation in every tensor-core precision mode; a shared (stride-0) source runs
    the encoder once; the source cache drops the encoder launches; u8 frames and the no-dense-motion constructor corner work."""
    from eamm_b200 import _lib as L
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    monkeypatch.setenv("EAMM_TC_NUM_SMS", "148")
    for k in ("EAMM_B200_MIX", "EAMM_B200_MIX_HG", "EAMM_B200_MIX_SKIP", "EAMM_B200_MIX64", "EAMM_B200_CONV", "EAMM_TC_ROW7"):
        monkeypatch.delenv(k, raising=False)
    ws = torch.zeros(L.SPLITK_WS_BYTES, dtype=torch.uint8)
    monkeypatch.setattr(engine, "current_stream_ptr", lambda: None)
    monkeypatch.setattr(engine, "splitk_workspace", lambda device, stream: ws)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)       # (asks the driver: none on this host)
    real, log = L.load(), []
    launches0 = L.LAUNCHES

    def build(cfg_name, precision):
        cfg = get_config(cfg_name)
        m = OcclusionAwareGenerator(**cfg).eval()
        m.load_state_dict(synth.make_state_dict(cfg, seed=0))
        e = engine.GeneratorEngine(m, precision)
        e.lib = _DryLib(real, log)
        if e.dm is not None:
            e.dm.lib = e.lib
        return m, e, cfg

    with torch.no_grad():
        for precision in ("fp32", "fp32_bf16x3", "fp16", "bf16"):
            m, e, cfg = build("full", precision)
            src, kpd, kps = synth.make_inputs(2, cfg, size=256, seed=1)
            del log[:]
            out = e.run(src, kpd, kps)
            assert [r[0] for r in log] == FRAME_LAUNCHES, precision
            assert all(r[1] == 0 for r in log), (precision, [x for x in log if x[1]])
            assert {k: tuple(v.shape) for k, v in out.items()} == {
                "mask": (2, 11, 64, 64), "sparse_deformed": (2, 11, 3, 64, 64), "occlusion_map": (2, 1, 64, 64),
                "deformed": (2, 3, 256, 256), "prediction": (2, 3, 256, 256)}
            assert list(out) == ["mask", "sparse_deformed", "occlusion_map", "deformed", "prediction"]    # generator.py:66-95 order
        # the bench configuration itself (BASELINE configs[1]: 32 frames, fp32 mode): the plans DESIGN.md section 4 describes
        m, e, cfg = build("full", "fp32")
        src, kpd, kps = synth.make_inputs(32, cfg, size=256, seed=1)
        del log[:]
        e.run(src, kpd, kps)
        assert [r[0] for r in log] == FRAME_LAUNCHES and all(r[1] == 0 for r in log)
        q = [r[3] for r in log if r[0] == "conv_tc"]            # (N tile, scheme, fold, chunks/stage, pair | halo<<1 | splitk<<8, stages)
        pair, halo, splitk = (lambda x: x[4] & 1), (lambda x: (x[4] >> 1) & 1), (lambda x: x[4] >> 8)
        first, down0, down1, hg, head, res, up0, up1, final = q[0], q[1], q[2], q[3:13], q[13], q[14:26], q[26], q[27], q[28]
        assert first[0] == 64 and first[2] == 2                                   # packed row-7, hi/lo folded along N
        assert (down0[0], pair(down0), halo(down0)) == (128, 1, 1) and (down1[0], pair(down1), halo(down1)) == (256, 1, 1)
        assert all((x[0], pair(x), halo(x), splitk(x)) == (256, 1, 1, 1) for x in res)            # CTA pairs, halo tiles
        assert (up0[0], halo(up0), up1[0], halo(up1)) == (128, 1, 64, 1)
        assert (head[1], head[0], final[1], final[0]) == (4, 112, 3, 112)                             # kx-in-N 7x7 schemes
        assert [splitk(x) for x in hg] == [1, 1, 1, 1, 9, 8, 4, 1, 1, 1]        # enc4 / dec0 / dec1: too few tiles for 148 SMs
        assert all(halo(x) == 0 for x in hg)                                                           # bf16 hi/lo: per-tap tiles
        del m, e, src
        # fp32 mode again (mixed formats): one shared source for 3 frames -> the encoder (3 convs) sees one image
        m, e, cfg = build("full", "fp32")
        src, kpd, kps = synth.make_inputs(3, cfg, size=256, seed=1, shared_source=True)
        del log[:]
        e.run(src[:1].expand(3, -1, -1, -1), kpd, kps)
        convs = [r[2] for r in log if r[0] == "conv_tc"]
        assert convs[:3] == [1, 1, 1] and set(convs[3:]) == {3} and all(r[1] == 0 for r in log)
        # source cache: the same source tensor again -> per-frame kernels only (no pack / first / down0 / down1 / anti-alias)
        m.cache_source = True
        one = src[:1].expand(3, -1, -1, -1)
        e.run(one, kpd, kps)
        del log[:]
        e.run(one, kpd, kps)
        assert [r[0] for r in log] == FRAME_LAUNCHES[5:] and len(log) == 30
        one2 = src[:1].clone().expand(3, -1, -1, -1)                    # another tensor: recomputed
        del log[:]
        e.run(one2, kpd, kps)
        assert len(log) == 35
        m.cache_source = False
        m.emit_u8 = True
        out = e.run(one2, kpd, kps)
        assert out["prediction_u8"].shape == (3, 256, 256, 3) and out["prediction_u8"].dtype == torch.uint8
        # constructor corner: no dense-motion network (generator.py:67) -> encoder, copy + norm1, bottleneck, decoder
        m, e, cfg = build("tiny_nodm", "fp32")
        src, kpd, kps = synth.make_inputs(2, cfg, size=64, seed=1)
        del log[:]
        out = e.run(src, kpd, kps)
        names = [r[0] for r in log]
        assert set(out) == {"prediction"} and "kp_stage" not in names and names.count("warp_occlude") == 1
        assert all(r[1] == 0 for r in log)
        # empty batch: nothing is launched, empty tensors of the reference's shapes come back
        m, e, cfg = build("full", "fp16")
        del log[:]
        src, kpd, kps = synth.make_inputs(1, cfg, size=256, seed=1)
        out = e.run(src[:0], {k: v[:0] for k, v in kpd.items()}, {k: v[:0] for k, v in kps.items()})
        assert not log and out["prediction"].shape == (0, 3, 256, 256) and out["mask"].shape == (0, 11, 64, 64)
    assert L.LAUNCHES > launches0                                        # (stubbed launches are counted like real ones)


def test_kp_detector_and_at_net2_host_programs_dry_run(monkeypatch):
    """The same dry run (launches stubbed, eamm_conv_tc -> its own argument check) for the SURVEY 8(f) engines: keypoint
    detectors in tensor-core modes and AT_net2 with its default and opt-in conv back ends; launch counts as documented."""
    from eamm_b200 import _lib as L, kp_engine, at_engine
    from eamm_b200.config import get_kp_config
    from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
    from eamm_b200.modules.util import AT_net2
    monkeypatch.setenv("EAMM_TC_NUM_SMS", "148")
    ws = torch.zeros(L.SPLITK_WS_BYTES, dtype=torch.uint8)
    for mod in (engine, kp_engine, at_engine):
        monkeypatch.setattr(mod, "current_stream_ptr", lambda: None)
    monkeypatch.setattr(engine, "splitk_workspace", lambda device, stream: ws)
    real, log = L.load(), []
    with torch.no_grad():
        for audio, cls in ((False, KPDetector), (True, KPDetector_a)):
            cfg = get_kp_config("full", audio=audio)
            m = cls(**cfg).eval()
            m.load_state_dict(synth.make_kp_state_dict(cfg))
            x = synth.make_kp_inputs(cfg, 2, 256, audio)
            for precision in ("fp32", "fp16", "bf16"):
                e = kp_engine.KPDetectorEngine(m, precision)
                e.lib = _DryLib(real, log)
                del log[:]
                out = e.run(x)
                names = [r[0] for r in log]
                assert all(r[1] == 0 for r in log), (audio, precision, [r for r in log if r[1]])
                # KPDetector: anti-alias, 10 Hourglass convs, merged head conv, head kernel; KPDetector_a: layout, conv, head
                assert names == (["nchw_to_act", "conv_tc", "kp_head"] if audio else
                                 ["aa_downsample_act"] + ["conv_tc"] * 11 + ["kp_head"]), (audio, precision, names)
                assert out["value"].shape == (2, 10, 2) and out["jacobian"].shape == (2, 10, 2, 2)
                assert out["heatmap"].shape == (2, 10, 58, 58)                    # valid 7x7 conv on the 64 x 64 map
        sd = synth.make_at_state_dict()
        img, mfcc, pose = synth.make_at_inputs(1, 3)
        for audio_tc, decon in (("simt", "tc"), ("tc", "tc"), ("simt", "simt")):
            monkeypatch.setenv("EAMM_B200_AT_AUDIO", audio_tc)
            monkeypatch.setenv("EAMM_B200_AT_DECON", decon)
            m = AT_net2().eval()
            m.load_state_dict(dict(sd), strict=True)
            e = at_engine.ATNet2Engine(m)
            e.lib = _DryLib(real, log)
            del log[:]
            out = e.run(img, mfcc, pose, 1.6)
            names = [r[0] for r in log]
            assert out.shape == (1, 3, 35, 64, 64) and all(r[1] == 0 for r in log), (audio_tc, decon)
            n_tc = (4 if audio_tc == "tc" else 0) + (4 if decon == "tc" else 0)
            assert names.count("conv_tc") == n_tc and names.count("conv_simt") == 8 + 5 + 4 - n_tc
            assert names.count("lstm_layer") == 3 and names.count("maxpool") == 2 and names.count("linear") == 9
            assert names.count("act_copy") == (6 if audio_tc == "tc" else 0)


def test_at_net2_engine_packing_reproduces_the_oracle_on_cpu():
    """Re-executes every packed stage of ATNet2Engine with torch (the kernels' documented semantics) and compares with
    the oracle: pins BN folding, the (c,h,w)->(h,w,c) FC permutation, the LSTM layer-0 split, the 1x1 ConvTranspose as
    a GEMM and ConvTranspose(k4,s2,p1) as UP2 parity classes."""
    from eamm_b200.modules.util import AT_net2
    from eamm_b200.at_engine import ATNet2Engine
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    sd = synth.make_at_state_dict()
    m = AT_net2().eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        eng = ATNet2Engine(m)
    B, T = 1, 2
    img, mfcc, pose = synth.make_at_inputs(B, T)
    taps = {}
    want = oracle.at_net2_forward(sd, img, mfcc, pose, 1.6, taps)
    lin = lambda l, x, scale=1.0, add=0: (F.relu(x @ l.w + (l.b if l.b is not None else 0) + add) if l.relu
                                         else x @ l.w + (l.b if l.b is not None else 0) + add) * scale
    x = img
    for lay in eng.img:
        x = _emulate_conv3(lay, x)
    img_feat = x.view(B, 512)
    a = mfcc.view(B * T, 1, 28, 12)
    a = _emulate_conv3(eng.aud[1], _emulate_conv3(eng.aud[0], a))
    a = F.max_pool2d(a, 3, stride=(1, 2))
    a = _emulate_conv3(eng.aud[5], _emulate_conv3(eng.aud[4], _emulate_conv3(eng.aud[3], a)))
    a = F.max_pool2d(a, 3, stride=(2, 2))
    flat = a.permute(0, 2, 3, 1).reshape(B * T, -1)                        # NHWC flatten, as the engine's buffer
    x2 = torch.cat([lin(eng.fc2, lin(eng.fc1, flat), 1.6), lin(eng.pose2, lin(eng.pose1, pose.view(B * T, 6)))], 1)
    want_in = taps["lstm_input"].view(B * T, -1)
    assert torch.allclose(x2, want_in[:, 512:], atol=2e-4)
    assert torch.allclose(img_feat, want_in[:1, :512], atol=1e-5)
    h = x2
    for l, ((p_img, p_x), whh) in enumerate(eng.lstm):
        add = lin(p_img, img_feat).repeat_interleave(T, 0) if p_img is not None else 0
        gates = lin(p_x, h, add=add).view(B, T, 1024)
        hs, hh, cc = [], torch.zeros(B, 256), torch.zeros(B, 256)
        for t in range(T):
            g = gates[:, t] + hh @ whh.t()
            cc = torch.sigmoid(g[:, 256:512]) * cc + torch.sigmoid(g[:, :256]) * torch.tanh(g[:, 512:768])
            hh = torch.sigmoid(g[:, 768:]) * torch.tanh(cc)
            hs.append(hh)
        h = torch.stack(hs, 1).view(B * T, 256)
    assert torch.allclose(h, taps["lstm_out"].view(B * T, 256), atol=1e-5)
    # (tensor-core decon stack: the 1x1 -> 4x4 GEMM writes NCHW, (c, y, x) columns; the SIMT variant writes NHWC)
    d = lin(eng.dec0, h).view(B * T, 256, 4, 4) if eng.dec_tc else lin(eng.dec0, h).view(B * T, 4, 4, 256).permute(0, 3, 1, 2)
    for lay in eng.dec:
        d = _emulate_up2(lay, d)
    got = d[:, :35].view(B, T, 35, 64, 64)
    assert torch.allclose(got, want, atol=2e-4), float((got - want).abs().max())
