"""Deterministic synthetic weights and inputs for the generation hot path.

The reference ships no checkpoints (SURVEY.md §2 notes), so parity and benchmarks run on seeded
random weights laid out exactly like ``OcclusionAwareGenerator.state_dict()`` of the reference
(196 keys for the full config; key names follow /root/reference/modules/generator.py:14-48,
dense_motion.py:12-30 and util.py:858-1052).  BatchNorm statistics and affine terms are randomised
(default init mean 0 / var 1 would hide BN-folding bugs).

Input recipe follows SURVEY.md §8(d): source ``rand(B,3,H,W)``, ``kp.value = rand*1.6-0.8``,
``kp.jacobian = eye(2) + 0.1*randn``.
"""
import math
from collections import OrderedDict

import torch


def conv_layers(cfg):
    """[(prefix, cin, cout, k, has_norm)] for every nn.Conv2d of the generator, reference key order."""
    nc, nkp = cfg["num_channels"], cfg["num_kp"]
    be, mf = cfg["block_expansion"], cfg["max_features"]
    out = []
    dm = cfg.get("dense_motion_params")
    if dm is not None:
        dbe, dmf, nb = dm["block_expansion"], dm["max_features"], dm["num_blocks"]
        cin0 = (nkp + 1) * (nc + 1)
        p = "dense_motion_network."
        for i in range(nb):                                   # util.py:948-952
            ci = cin0 if i == 0 else min(dmf, dbe * 2 ** i)
            co = min(dmf, dbe * 2 ** (i + 1))
            out.append((p + f"hourglass.encoder.down_blocks.{i}", ci, co, 3, True))
        for j, i in enumerate(range(nb)[::-1]):               # util.py:973-977
            ci = (1 if i == nb - 1 else 2) * min(dmf, dbe * 2 ** (i + 1))
            co = min(dmf, dbe * 2 ** i)
            out.append((p + f"hourglass.decoder.up_blocks.{j}", ci, co, 3, True))
        hg_out = dbe + cin0                                    # util.py:980
        out.append((p + "mask", hg_out, nkp + 1, 7, False))
        if cfg.get("estimate_occlusion_map", False):
            out.append((p + "occlusion", hg_out, 1, 7, False))
    out.append(("first", nc, be, 7, True))
    nd = cfg["num_down_blocks"]
    for i in range(nd):
        out.append((f"down_blocks.{i}", min(mf, be * 2 ** i), min(mf, be * 2 ** (i + 1)), 3, True))
    for i in range(nd):
        out.append((f"up_blocks.{i}", min(mf, be * 2 ** (nd - i)), min(mf, be * 2 ** (nd - i - 1)), 3, True))
    cb = min(mf, be * 2 ** nd)
    for i in range(cfg["num_bottleneck_blocks"]):
        out.append((f"bottleneck.r{i}", cb, cb, 3, "res"))
    out.append(("final", be, nc, 7, False))
    return out


def _bn(sd, name, c, g):
    sd[name + ".weight"] = torch.rand(c, generator=g) + 0.5
    sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
    sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
    sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5
    sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)


def _conv(sd, name, cin, cout, k, g, gain=1.0):
    bound = gain * math.sqrt(3.0 / (cin * k * k))              # unit-gain fan-in uniform
    sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
    sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * 0.1


def make_state_dict(cfg, seed=0):
    """Seeded fp32 CPU state dict with the reference's key layout."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    dm = cfg.get("dense_motion_params")
    for prefix, cin, cout, k, kind in conv_layers(cfg):
        if kind == "res":                                       # util.py:863-870
            _conv(sd, prefix + ".conv1", cin, cout, k, g, gain=1.4)
            _conv(sd, prefix + ".conv2", cin, cout, k, g, gain=0.7)
            _bn(sd, prefix + ".norm1", cin, g)
            _bn(sd, prefix + ".norm2", cin, g)
        elif kind:
            _conv(sd, prefix + ".conv", cin, cout, k, g, gain=1.4)
            _bn(sd, prefix + ".norm", cout, g)
        else:
            _conv(sd, prefix, cin, cout, k, g, gain=0.5 if prefix == "final" else 2.0)
    if dm is not None and dm.get("scale_factor", 1) != 1:       # dense_motion.py:29-30 buffer
        sd["dense_motion_network.down.weight"] = aa_kernel(cfg["num_channels"])
    return sd


def make_kp_state_dict(cfg, seed=2):
    """Seeded state dict laid out like KPDetector / KPDetector_a (keypoint_detector.py:14-37, 117-140):
    `predictor` Hourglass, `kp` and `jacobian` 7x7 convs, `down.weight` buffer."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    be, mf, nb = cfg["block_expansion"], cfg["max_features"], cfg["num_blocks"]
    cin0 = cfg.get("num_channels_a", cfg["num_channels"])
    for i in range(nb):
        ci = cin0 if i == 0 else min(mf, be * 2 ** i)
        co = min(mf, be * 2 ** (i + 1))
        _conv(sd, f"predictor.encoder.down_blocks.{i}.conv", ci, co, 3, g, gain=1.4)
        _bn(sd, f"predictor.encoder.down_blocks.{i}.norm", co, g)
    for j, i in enumerate(range(nb)[::-1]):
        ci = (1 if i == nb - 1 else 2) * min(mf, be * 2 ** (i + 1))
        co = min(mf, be * 2 ** i)
        _conv(sd, f"predictor.decoder.up_blocks.{j}.conv", ci, co, 3, g, gain=1.4)
        _bn(sd, f"predictor.decoder.up_blocks.{j}.norm", co, g)
    feat = be + cin0
    _conv(sd, "kp", feat, cfg["num_kp"], 7, g, gain=1.0)
    if cfg.get("estimate_jacobian", False):
        maps = 1 if cfg.get("single_jacobian_map", False) else cfg["num_kp"]
        _conv(sd, "jacobian", feat, 4 * maps, 7, g, gain=0.5)
        sd["jacobian.bias"] = sd["jacobian.bias"] + torch.tensor([1.0, 0.0, 0.0, 1.0] * maps)
    if cfg.get("scale_factor", 1) != 1:
        sd["down.weight"] = aa_kernel(cfg["num_channels"])
    return sd


def make_kp_inputs(cfg, batch, size, audio, seed=5):
    """Seeded input of the keypoint detectors: an image for KPDetector, a feature map for KPDetector_a."""
    g = torch.Generator().manual_seed(seed)
    if audio:
        step = int(1 / cfg["scale_factor"])
        return torch.randn(batch, cfg["block_expansion"] + cfg["num_channels_a"], size // step, size // step, generator=g)
    return torch.rand(batch, cfg["num_channels"], size, size, generator=g)


def make_clip_inputs(T=12, K=10, Ke=4, seed=7):
    """Seeded per-frame detector outputs of a clip: driving kp [T,K,...], emotion kp [T,Ke,...], source and
    initial-driving keypoints (batch 1).  Shapes follow demo.py:206-228."""
    g = torch.Generator().manual_seed(seed)
    eye = torch.eye(2).view(1, 1, 2, 2)
    drv = {"value": torch.rand(T, K, 2, generator=g) * 1.2 - 0.6, "jacobian": eye + 0.1 * torch.randn(T, K, 2, 2, generator=g)}
    emo = {"value": torch.randn(T, Ke, 2, generator=g) * 0.05, "jacobian": 0.05 * torch.randn(T, Ke, 2, 2, generator=g)}
    src = {"value": torch.rand(1, K, 2, generator=g) * 1.2 - 0.6, "jacobian": eye + 0.1 * torch.randn(1, K, 2, 2, generator=g)}
    init = {"value": drv["value"][:1].clone(), "jacobian": drv["jacobian"][:1].clone()}
    return drv, emo, src, init


def _linear(sd, name, cin, cout, g, gain=1.4):
    sd[name + ".weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * (gain * math.sqrt(3.0 / cin))
    sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * 0.1


def make_at_state_dict(seed=4):
    """Seeded state dict laid out like AT_net2 (util.py:514-577) minus its unused StyleGAN2 `generator.*` branch
    (only reached with jaco_net == 'gan'; every shipped config uses 'cnn').  35.3 M parameters."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for i in range(8):                                          # util.py:518-522
        ci, co = (3 if i == 0 else 2 * 2 ** i), 2 * 2 ** (i + 1)
        _conv(sd, f"down_blocks.{i}.conv", ci, co, 3, g, gain=1.6)
        _bn(sd, f"down_blocks.{i}.norm", co, g)
    _linear(sd, "pose_encoder.0", 6, 128, g)
    _linear(sd, "pose_encoder.2", 128, 256, g)
    for i, (ci, co) in zip((0, 1, 3, 4, 5), ((1, 64), (64, 128), (128, 256), (256, 256), (256, 512))):
        _conv(sd, f"audio_eocder.{i}.0", ci, co, 3, g, gain=1.6)
        del sd[f"audio_eocder.{i}.0.bias"]                      # util.py:1745: no bias under a normaliser
        _bn(sd, f"audio_eocder.{i}.1", co, g)
    _linear(sd, "audio_eocder_fc.0", 1024 * 12, 2048, g)
    _linear(sd, "audio_eocder_fc.2", 2048, 256, g)
    for l in range(3):                                          # nn.LSTM(1024, 256, 3); wider than the default U(-1/16, 1/16) so h is not tiny
        for nm, shape in (("weight_ih", (1024, 1024 if l == 0 else 256)), ("weight_hh", (1024, 256)),
                          ("bias_ih", (1024,)), ("bias_hh", (1024,))):
            sd[f"lstm.{nm}_l{l}"] = (torch.rand(*shape, generator=g) * 2 - 1) / (16.0 if shape[-1] == 1024 else 5.0)
    for i, (ci, co, k) in zip((0, 3, 6, 9, 12), ((256, 256, 6), (256, 128, 4), (128, 128, 4), (128, 128, 4), (128, 35, 4))):
        taps = 1 if k == 6 else 4                              # input taps that reach one output pixel
        sd[f"decon.{i}.weight"] = (torch.rand(ci, co, k, k, generator=g) * 2 - 1) * ((0.8 if i == 12 else 1.6) * math.sqrt(3.0 / (ci * taps)))
        sd[f"decon.{i}.bias"] = (torch.rand(co, generator=g) * 2 - 1) * 0.1
        if i != 12:
            _bn(sd, f"decon.{i + 1}", co, g)
    return sd


def make_at_inputs(B, T, seed=6):
    """(example_image [B,3,256,256], mfcc windows [B,T,28,12], pose [B,T,6]); value ranges of demo.py:316-343."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, 256, 256, generator=g)
    mfcc = torch.randn(B, T, 28, 12, generator=g) * 0.5
    pose = torch.randn(B, T, 6, generator=g) * 0.3
    return img, mfcc, pose


def aa_kernel(channels, sigma=1.5):
    """The fixed 13x13 Gaussian buffer of AntiAliasInterpolation2d (util.py:1012-1036)."""
    ks = 2 * round(sigma * 4) + 1
    ax = torch.arange(ks, dtype=torch.float32)
    mean = (ks - 1) / 2
    g1 = torch.exp(-(ax - mean) ** 2 / (2 * sigma ** 2))
    k2 = g1[:, None] * g1[None, :]
    k2 = k2 / torch.sum(k2)
    return k2.view(1, 1, ks, ks).repeat(channels, 1, 1, 1)


def make_inputs(batch, cfg, size=256, seed=1, shared_source=False, with_jacobian=True):
    """(source_image, kp_driving, kp_source) fp32 CPU tensors, SURVEY.md §8(d) recipe."""
    g = torch.Generator().manual_seed(seed)
    nkp, nc = cfg["num_kp"], cfg["num_channels"]
    nsrc = 1 if shared_source else batch
    src = torch.rand(nsrc, nc, size, size, generator=g)
    kps = {"value": torch.rand(nsrc, nkp, 2, generator=g) * 1.6 - 0.8}
    kpd = {"value": torch.rand(batch, nkp, 2, generator=g) * 1.6 - 0.8}
    if with_jacobian:
        eye = torch.eye(2).view(1, 1, 2, 2)
        kps["jacobian"] = eye + 0.1 * torch.randn(nsrc, nkp, 2, 2, generator=g)
        kpd["jacobian"] = eye + 0.1 * torch.randn(batch, nkp, 2, 2, generator=g)
    if shared_source:
        src = src.expand(batch, -1, -1, -1).contiguous()
        kps = {k: v.expand(batch, *v.shape[1:]).contiguous() for k, v in kps.items()}
    return src, kpd, kps
