"""CPU ORACLE for the EAMM per-frame generation hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a ``state_dict``, what the reference computes in
``OcclusionAwareGenerator.forward`` (/root/reference/modules/generator.py:59-97) and
``DenseMotionNetwork.forward`` (/root/reference/modules/dense_motion.py:81-113) with the building
blocks of /root/reference/modules/util.py:815-1052 and the eval branch of
/root/reference/sync_batchnorm/batchnorm.py:48-53.

Where the arithmetic lives: the reference delegates every numeric op to PyTorch (pinned
torch==1.10.1 in /root/reference/requirements.txt:1; the container has torch 2.11.0 whose defaults
for these ops are identical: grid_sample bilinear/zeros/align_corners=False, interpolate nearest,
bilinear interpolate align_corners=False, BN eps 1e-5).  The oracle therefore calls the same
``torch.nn.functional`` CPU ops at the same call sites; ``oracle/sampler_np.py`` additionally
restates the sampling/interpolation index arithmetic in numpy so the integer tap selection is
pinned independently of torch.

PARITY PIN: the reference ships no tests, golden vectors or checkpoints, so the
reference's own tests pin nothing.  The pin used here is the reference itself: with
/root/reference importable (authoring container only) ``tools/make_golden.py`` loads the same
seeded state dict into the real ``OcclusionAwareGenerator`` and checks this oracle against it
bit-for-bit, then writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the
oracle against those fixtures everywhere else.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs (cpu_baseline, ``--impl reference``,
and cudnn_baseline = these same functions on CUDA tensors as the eager-PyTorch bar; none of them is the thing
measured as ours) may import this module.  The product path (``eamm_b200``) never does.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # sync_batchnorm/batchnorm.py:39 default eps


# ----------------------------------------------------------------------------- building blocks
def batch_norm_eval(x, sd, p):
    """sync_batchnorm/batchnorm.py:50-53 with training=False."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.1, BN_EPS)


def same_block(x, sd, p, pad):
    """SameBlock2d.forward, util.py:934-938."""
    out = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=pad)
    return F.relu(batch_norm_eval(out, sd, p + ".norm"))


def down_block(x, sd, p):
    """DownBlock2d.forward, util.py:915-920 (conv3x3 pad1 -> BN -> ReLU -> AvgPool 2x2)."""
    out = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
    out = F.relu(batch_norm_eval(out, sd, p + ".norm"))
    return F.avg_pool2d(out, kernel_size=(2, 2))


def up_block(x, sd, p):
    """UpBlock2d.forward, util.py:895-900 (nearest x2 -> conv3x3 pad1 -> BN -> ReLU)."""
    out = F.interpolate(x, scale_factor=2)
    out = F.conv2d(out, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
    return F.relu(batch_norm_eval(out, sd, p + ".norm"))


def res_block(x, sd, p):
    """ResBlock2d.forward, util.py:872-880 (pre-activation)."""
    out = F.relu(batch_norm_eval(x, sd, p + ".norm1"))
    out = F.conv2d(out, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    out = F.relu(batch_norm_eval(out, sd, p + ".norm2"))
    out = F.conv2d(out, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    return out + x


def make_coordinate_grid(h, w, device=None):
    """util.py:839-855: x = 2*i/(w-1)-1, y = 2*j/(h-1)-1, last dim (x, y).  (`device`: the reference builds the grid
    with `.type(kp.type())`, i.e. on the keypoints' device; bench.py's eager-CUDA baseline leg uses that.)"""
    x = torch.arange(w, device=device).type(torch.float32)
    y = torch.arange(h, device=device).type(torch.float32)
    x = (2 * (x / (w - 1)) - 1)
    y = (2 * (y / (h - 1)) - 1)
    yy = y.view(-1, 1).repeat(1, w)
    xx = x.view(1, -1).repeat(h, 1)
    return torch.cat([xx.unsqueeze(2), yy.unsqueeze(2)], 2)


def kp2gaussian(value, h, w, kp_variance):
    """util.py:815-836."""
    grid = make_coordinate_grid(h, w, value.device).view(1, 1, h, w, 2)
    mean_sub = grid - value.view(value.shape[0], value.shape[1], 1, 1, 2)
    return torch.exp(-0.5 * (mean_sub ** 2).sum(-1) / kp_variance)


def anti_alias_down(x, weight, scale):
    """AntiAliasInterpolation2d.forward, util.py:1044-1052."""
    if scale == 1.0:
        return x
    ks = weight.shape[-1]
    ka = ks // 2
    kb = ka - 1 if ks % 2 == 0 else ka
    out = F.pad(x, (ka, kb, ka, kb))
    out = F.conv2d(out, weight=weight, groups=x.shape[1])
    s = int(1 / scale)
    return out[:, :, ::s, ::s]


# ----------------------------------------------------------------------------- dense motion
def heatmap_representation(kp_driving, kp_source, h, w, kp_variance):
    """create_heatmap_representations, dense_motion.py:32-45 -> [B, K+1, 1, h, w]."""
    hm = kp2gaussian(kp_driving["value"], h, w, kp_variance) - kp2gaussian(kp_source["value"], h, w, kp_variance)
    zeros = torch.zeros(hm.shape[0], 1, h, w, dtype=hm.dtype, device=hm.device)
    return torch.cat([zeros, hm], dim=1).unsqueeze(2)


def sparse_motions(kp_driving, kp_source, h, w):
    """create_sparse_motions, dense_motion.py:47-67 -> [B, K+1, h, w, 2]."""
    bs, nkp = kp_driving["value"].shape[:2]
    identity = make_coordinate_grid(h, w, kp_driving["value"].device).view(1, 1, h, w, 2)
    grid = identity - kp_driving["value"].view(bs, nkp, 1, 1, 2)
    if "jacobian" in kp_driving:
        jac = torch.matmul(kp_source["jacobian"], torch.inverse(kp_driving["jacobian"]))
        jac = jac.unsqueeze(-3).unsqueeze(-3).repeat(1, 1, h, w, 1, 1)
        grid = torch.matmul(jac, grid.unsqueeze(-1)).squeeze(-1)
    d2s = grid + kp_source["value"].view(bs, nkp, 1, 1, 2)
    return torch.cat([identity.repeat(bs, 1, 1, 1, 1), d2s], dim=1)


def deformed_source(src_small, sparse_motion):
    """create_deformed_source_image, dense_motion.py:69-79 -> [B, K+1, C, h, w]."""
    bs, c, h, w = src_small.shape
    k1 = sparse_motion.shape[1]
    rep = src_small.unsqueeze(1).unsqueeze(1).repeat(1, k1, 1, 1, 1, 1).view(bs * k1, -1, h, w)
    out = F.grid_sample(rep, sparse_motion.reshape(bs * k1, h, w, -1), align_corners=False)
    return out.view(bs, k1, -1, h, w)


def hourglass(x, sd, p, num_blocks):
    """Hourglass.forward = Decoder(Encoder(x)), util.py:955-1002."""
    outs = [x]
    for i in range(num_blocks):
        outs.append(down_block(outs[-1], sd, f"{p}.encoder.down_blocks.{i}"))
    out = outs.pop()
    for j in range(num_blocks):
        out = up_block(out, sd, f"{p}.decoder.up_blocks.{j}")
        out = torch.cat([out, outs.pop()], dim=1)
    return out


def dense_motion_forward(sd, cfg, source_image, kp_driving, kp_source, prefix="dense_motion_network",
                         taps=None):
    """DenseMotionNetwork.forward, dense_motion.py:81-113."""
    dm = cfg["dense_motion_params"]
    scale = dm.get("scale_factor", 1)
    kp_var = dm.get("kp_variance", 0.01)
    if scale != 1:
        source_image = anti_alias_down(source_image, sd[prefix + ".down.weight"], scale)
    bs, _, h, w = source_image.shape
    out = {}
    hm = heatmap_representation(kp_driving, kp_source, h, w, kp_var)
    sm = sparse_motions(kp_driving, kp_source, h, w)
    ds = deformed_source(source_image, sm)
    out["sparse_deformed"] = ds
    inp = torch.cat([hm, ds], dim=2).view(bs, -1, h, w)
    pred = hourglass(inp, sd, prefix + ".hourglass", dm["num_blocks"])
    mask = F.softmax(F.conv2d(pred, sd[prefix + ".mask.weight"], sd[prefix + ".mask.bias"], padding=3), dim=1)
    out["mask"] = mask
    deformation = (sm.permute(0, 1, 4, 2, 3) * mask.unsqueeze(2)).sum(dim=1).permute(0, 2, 3, 1)
    out["deformation"] = deformation
    if cfg.get("estimate_occlusion_map", False):
        out["occlusion_map"] = torch.sigmoid(
            F.conv2d(pred, sd[prefix + ".occlusion.weight"], sd[prefix + ".occlusion.bias"], padding=3))
    if taps is not None:
        taps.update(source_small=source_image, heatmap=hm, sparse_motion=sm, hourglass_in=inp, hourglass_out=pred)
    return out


# ----------------------------------------------------------------------------- generator
def deform_input(inp, deformation):
    """OcclusionAwareGenerator.deform_input, generator.py:50-57."""
    _, h_old, w_old, _ = deformation.shape
    _, _, h, w = inp.shape
    if h_old != h or w_old != w:
        deformation = deformation.permute(0, 3, 1, 2)
        deformation = F.interpolate(deformation, size=(h, w), mode="bilinear", align_corners=False)
        deformation = deformation.permute(0, 2, 3, 1)
    return F.grid_sample(inp, deformation, align_corners=False)


def generator_forward(sd, cfg, source_image, kp_driving, kp_source, taps=None):
    """OcclusionAwareGenerator.forward, generator.py:59-97.  Returns the same dict of tensors."""
    with torch.no_grad():
        out = same_block(source_image, sd, "first", 3)
        if taps is not None:
            taps["first"] = out
        for i in range(cfg["num_down_blocks"]):
            out = down_block(out, sd, f"down_blocks.{i}")
        if taps is not None:
            taps["encoded"] = out
        result = {}
        if cfg.get("dense_motion_params") is not None:
            dmo = dense_motion_forward(sd, cfg, source_image, kp_driving, kp_source, taps=taps)
            result["mask"] = dmo["mask"]
            result["sparse_deformed"] = dmo["sparse_deformed"]
            occ = dmo.get("occlusion_map")
            if occ is not None:
                result["occlusion_map"] = occ
            deformation = dmo["deformation"]
            out = deform_input(out, deformation)
            if occ is not None:
                if out.shape[2] != occ.shape[2] or out.shape[3] != occ.shape[3]:
                    occ = F.interpolate(occ, size=out.shape[2:], mode="bilinear", align_corners=False)
                out = out * occ
            result["deformed"] = deform_input(source_image, deformation)
            if taps is not None:
                taps["deformation"] = deformation
                taps["warped"] = out
        for i in range(cfg["num_bottleneck_blocks"]):
            out = res_block(out, sd, f"bottleneck.r{i}")
        if taps is not None:
            taps["bottleneck"] = out
        for i in range(cfg["num_down_blocks"]):
            out = up_block(out, sd, f"up_blocks.{i}")
        if taps is not None:
            taps["decoded"] = out
        out = F.conv2d(out, sd["final.weight"], sd["final.bias"], padding=3)
        result["prediction"] = torch.sigmoid(out)
    return result


# ----------------------------------------------------------------------------- keypoint detector heads
# the step immediately before the generation path (demo.py:206-219).
def gaussian2kp(heatmap):
    """KPDetector.gaussian2kp, keypoint_detector.py:40-50: value = sum(heatmap * grid) over (h, w)."""
    h, w = heatmap.shape[2:]
    grid = make_coordinate_grid(h, w).unsqueeze(0).unsqueeze(0)
    return (heatmap.unsqueeze(-1) * grid).sum(dim=(2, 3))


def kp_heads(feature_map, sd, cfg):
    """The tail shared by KPDetector.forward (keypoint_detector.py:82-103) and KPDetector_a.forward (:183-203)."""
    pad = cfg.get("pad", 0)
    prediction = F.conv2d(feature_map, sd["kp.weight"], sd["kp.bias"], padding=pad)
    shape = prediction.shape
    heatmap = F.softmax(prediction.view(shape[0], shape[1], -1) / cfg["temperature"], dim=2).view(*shape)
    out = {"value": gaussian2kp(heatmap), "heatmap": heatmap}
    if cfg.get("estimate_jacobian", False):
        maps = 1 if cfg.get("single_jacobian_map", False) else cfg["num_kp"]
        jm = F.conv2d(feature_map, sd["jacobian.weight"], sd["jacobian.bias"], padding=pad)
        jm = jm.reshape(shape[0], maps, 4, shape[2], shape[3])
        jac = (heatmap.unsqueeze(2) * jm).view(shape[0], shape[1], 4, -1).sum(dim=-1)
        out["jacobian"] = jac.view(shape[0], shape[1], 2, 2)
    return out


def kp_detector_forward(sd, cfg, x, taps=None):
    """KPDetector.forward, keypoint_detector.py:77-105: anti-alias down, Hourglass predictor, heads."""
    with torch.no_grad():
        if cfg.get("scale_factor", 1) != 1:
            x = anti_alias_down(x, sd["down.weight"], cfg["scale_factor"])
        feature_map = hourglass(x, sd, "predictor", cfg["num_blocks"])
        if taps is not None:
            taps["feature_map"] = feature_map
        return kp_heads(feature_map, sd, cfg)


def kp_detector_a_forward(sd, cfg, feature_map):
    """KPDetector_a.forward, keypoint_detector.py:180-205: heads only (its predictor is never called)."""
    with torch.no_grad():
        return kp_heads(feature_map, sd, cfg)


def frames_u8(prediction):
    """What demo.py does with `out['prediction']` before writing the video: :281 NCHW -> NHWC, :507
    `skimage.img_as_ubyte` (scikit-image is an unpinned, un-vendored dependency, requirements.txt:9; absent
    here).  Published semantics of skimage.util.dtype._convert for a float32 image in [-1, 1] -> uint8:
    multiply by 255 in float32, np.rint (round half to even), clip to [0, 255], cast."""
    import numpy as np
    a = prediction.permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32)
    return torch.from_numpy(np.clip(np.rint(a * np.float32(255.0)), 0, 255).astype(np.uint8))


# ----------------------------------------------------------------------------- AT_net2
def _bn2d(x, sd, p):
    """nn.BatchNorm2d in eval mode (eps 1e-5), used by util.py:1757 (`conv2d` helper) and the `decon` stack."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, 1e-5)


def at_audio_encoder(x, sd):
    """AT_net2.audio_eocder, util.py:540-548: five conv3x3(no bias)+BN+ReLU (util.py:1740-1753) and two
    MaxPool2d(3) with strides (1,2) and (2,2).  x [B,1,28,12] -> [B,512,12,2]."""
    def cbr(x, i):
        p = "audio_eocder.%d" % i
        return F.relu(_bn2d(F.conv2d(x, sd[p + ".0.weight"], None, padding=1), sd, p + ".1"))
    x = cbr(cbr(x, 0), 1)
    x = F.max_pool2d(x, 3, stride=(1, 2))
    x = cbr(cbr(cbr(x, 3), 4), 5)
    return F.max_pool2d(x, 3, stride=(2, 2))


def at_decon(z, sd):
    """AT_net2.decon, util.py:559-575: ConvTranspose2d(256,256,6,2,1) and four ConvTranspose2d(.,.,4,2,1), BN+ReLU
    between them.  z [B,256,1,1] -> [B,35,64,64]."""
    for i in (0, 3, 6, 9):
        z = F.conv_transpose2d(z, sd["decon.%d.weight" % i], sd["decon.%d.bias" % i], stride=2, padding=1)
        z = F.relu(_bn2d(z, sd, "decon.%d" % (i + 1)))
    return F.conv_transpose2d(z, sd["decon.12.weight"], sd["decon.12.bias"], stride=2, padding=1)


def lstm_explicit(x, sd, layers=3):
    """The recurrence nn.LSTM documents (gate order i, f, g, o; c' = f*c + i*g; h' = o*tanh(c')), zero initial
    state, batch_first.  Restated step by step; `at_net2_forward` itself calls torch's fused op, and the tests
    hold the two within float rounding of each other."""
    B, T, _ = x.shape
    for l in range(layers):
        wi, wh = sd["lstm.weight_ih_l%d" % l], sd["lstm.weight_hh_l%d" % l]
        b = sd["lstm.bias_ih_l%d" % l] + sd["lstm.bias_hh_l%d" % l]
        H = wh.shape[1]
        h, c = torch.zeros(B, H), torch.zeros(B, H)
        outs = []
        for t in range(T):
            g = x[:, t] @ wi.t() + h @ wh.t() + b
            i_, f_, g_, o_ = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
            c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
            h = torch.sigmoid(o_) * torch.tanh(c)
            outs.append(h)
        x = torch.stack(outs, 1)
    return x


def at_net2_forward(sd, example_image, audio, pose, weight, taps=None):
    """AT_net2.forward with jaco_net == 'cnn', util.py:580-613.
    example_image [B,3,256,256], audio [B,T,28,12] (MFCC windows), pose [B,T,6] -> [B,T,35,64,64]."""
    with torch.no_grad():
        B, T = audio.shape[:2]
        outs = example_image
        for i in range(8):                                              # util.py:583-586
            outs = down_block(outs, sd, "down_blocks.%d" % i)
        image_feature = outs.view(B, -1)
        lstm_input = []
        for t in range(T):                                              # util.py:588-595
            cur = at_audio_encoder(audio[:, t].unsqueeze(1), sd)
            cur = cur.view(B, -1)
            cur = F.relu(F.linear(cur, sd["audio_eocder_fc.0.weight"], sd["audio_eocder_fc.0.bias"]))
            cur = F.relu(F.linear(cur, sd["audio_eocder_fc.2.weight"], sd["audio_eocder_fc.2.bias"])) * weight
            p = F.relu(F.linear(pose[:, t], sd["pose_encoder.0.weight"], sd["pose_encoder.0.bias"]))
            p = F.relu(F.linear(p, sd["pose_encoder.2.weight"], sd["pose_encoder.2.bias"]))
            lstm_input.append(torch.cat([image_feature, cur, p], 1))
        lstm_input = torch.stack(lstm_input, dim=1)
        flat = []
        for l in range(3):
            flat += [sd["lstm.weight_ih_l%d" % l], sd["lstm.weight_hh_l%d" % l],
                     sd["lstm.bias_ih_l%d" % l], sd["lstm.bias_hh_l%d" % l]]
        hx = (torch.zeros(3, B, 256), torch.zeros(3, B, 256))           # util.py:581-582
        lstm_out = torch._VF.lstm(lstm_input, hx, flat, True, 3, 0.0, False, False, True)[0]   # nn.LSTM.forward
        if taps is not None:
            taps["lstm_input"], taps["lstm_out"] = lstm_input, lstm_out
        deco = [at_decon(lstm_out[:, t, :].unsqueeze(2).unsqueeze(3), sd) for t in range(T)]   # util.py:600-606
        return torch.stack(deco, dim=1)
