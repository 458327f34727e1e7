"""Executor of AT_net2, the per-clip audio -> motion-feature network (SURVEY.md section 8(f) rank 4;
reference /root/reference/modules/util.py:514-613, called once per clip from demo.py:345).

The whole clip is processed as one batch of B*T frames (the reference loops over T, util.py:588-606):
  image DownBlocks x8 (once per clip)       eamm_conv_simt 3x3 + folded BN + ReLU + 2x2 avg-pool
  MFCC encoder (5 conv + 2 max-pool)        eamm_conv_simt / eamm_maxpool on [B*T, 28, 12] maps
  audio FC, pose MLP                        eamm_linear (ReLU, x`weight` fused)
  3-layer LSTM                              eamm_linear (input projections for all t) + eamm_lstm_layer
  decon: ConvTranspose 1x1->4x4             eamm_linear (the 16 output pixels are 16 column blocks)
         4 x ConvTranspose(k4, s2, p1)      UP2 kind with parity-class weights, NCHW fp32 out: eamm_conv_tc on bf16 hi/lo
                                            planes (three tensor-core passes, fp32-equivalent: the maps are powers of two),
                                            or eamm_conv_simt with EAMM_B200_AT_DECON=simt
This stage runs once per clip and feeds the keypoint softmax (temperature 0.1), so it is kept at fp32(-equivalent)
accuracy in every precision mode of the frame path.  The MFCC convs stay on fp32 CUDA cores by default: they sit in front
of FC1 -> LSTM -> decon -> softmax(T = 0.1) and the chain amplifies their rounding ~100x.  Opt-in (EAMM_B200_AT_AUDIO=tc):
audio1 / 3 / 4 / 5 (28x12 and 26x5 maps, 93 % of the encoder's FLOPs) on tensor cores over maps padded to 32x16 / 32x8
(bf16 hi/lo planes, three passes; eamm_act_copy moves the activations into and out of the padded buffers and re-zeroes the
padding a convolution has written, so every layer still sees the reference's zero padding; the two max-pools and audio0
work on the exact maps).  Measured on B200, 300-frame LRW clip: AT_net2 10.2 -> 4.6 ms per clip, AT_net2 output still
within 1e-4 of the reference, but the normalised keypoints move by more than the 1e-3 the clip test states (worst frame
PSNR 57.3 -> 51.4 dB): 16 significant bits per operand are not enough in front of that chain.
"""
import os

import ctypes as C

import torch

from . import _lib as L
from .engine import ActBuf, ConvLayer, WorkspaceCache, _launch, _round_up, current_stream_ptr, fold_bn, bn_affine


def _bn(sd_mod):
    return {"weight": sd_mod.weight.detach().float(), "bias": sd_mod.bias.detach().float(),
            "running_mean": sd_mod.running_mean.detach().float(), "running_var": sd_mod.running_var.detach().float()}


def convT_parity_weights(w):
    """ConvTranspose2d(k=4, s=2, p=1) as four 2x2 convs on the input grid -- the UP2 kernel's shape.

    Output pixel (2y+a, 2x+b), tap (ty,tx) reads input (y+a-1+ty, x+b-1+tx) with kernel element
    (3-a-2ty, 3-b-2tx).  w [cin][cout][4][4] -> [4 classes][4 taps][cout][cin].
    """
    classes = []
    for a in (0, 1):
        for b in (0, 1):
            taps = [w[:, :, 3 - a - 2 * ty, 3 - b - 2 * tx].t() for ty in (0, 1) for tx in (0, 1)]
            classes.append(torch.stack(taps, 0))
    return torch.stack(classes, 0)


class Linear:
    def __init__(self, name, w, b, relu, dev):
        """w [N][K] as nn.Linear stores it."""
        self.name, self.relu = name, int(relu)
        self.N, self.K = w.shape
        self.w = w.t().contiguous().to(dev)
        self.b = None if b is None else b.contiguous().to(dev)

    def launch(self, lib, st, x, ldx, y, ldy, M, scale=1.0, add=None, period=1):
        _launch("linear:" + self.name, lambda: L.check(lib.eamm_linear(
            x, ldx, self.w.data_ptr(), None if self.b is None else self.b.data_ptr(),
            None if add is None else add.data_ptr(), period, y, ldy, M, self.K, self.N, self.relu, float(scale), st),
            "linear " + self.name), flops=2.0 * M * self.K * self.N)


class ATNet2Engine:
    HID = 256

    def __init__(self, module, precision=None):
        m = self.m = module
        self.lib = L.load()
        dev = self.device = m.lstm.weight_hh_l0.device
        f = lambda t: t.detach().float()
        # image branch: DownBlock2d x8 (util.py:518-522)
        self.img = []
        for i, blk in enumerate(m.down_blocks):
            w, b = fold_bn(f(blk.conv.weight), f(blk.conv.bias), _bn(blk.norm))
            self.img.append(ConvLayer("at.down%d" % i, L.CONV_3X3, L.EPI_RELU | L.EPI_POOL2, w, b,
                                      _round_up(w.shape[1], 4), 4, "simt"))
        # MFCC encoder (util.py:540-548)
        self.aud = {}
        self.aud_tc = os.environ.get("EAMM_B200_AT_AUDIO", "simt") == "tc"
        for i in (0, 1, 3, 4, 5):
            conv, norm = m.audio_eocder[i][0], m.audio_eocder[i][1]
            w, b = fold_bn(f(conv.weight), torch.zeros(conv.out_channels, device=dev), _bn(norm))
            tc = self.aud_tc and i != 0
            self.aud[i] = ConvLayer("at.audio%d" % i, L.CONV_3X3, L.EPI_RELU, w, b, _round_up(w.shape[1], 64 if tc else 4),
                                    16 if tc else 4, "tc3" if tc else "simt")
        # audio FC: the encoder output is flattened (c, h, w) in the reference, (h, w, c) here
        w1 = f(m.audio_eocder_fc[0].weight)
        hw = w1.shape[1] // 512
        w1 = w1.view(-1, 512, hw).permute(0, 2, 1).reshape(w1.shape[0], -1)
        self.fc1 = Linear("at.fc1", w1, f(m.audio_eocder_fc[0].bias), True, dev)
        self.fc2 = Linear("at.fc2", f(m.audio_eocder_fc[2].weight), f(m.audio_eocder_fc[2].bias), True, dev)
        self.pose1 = Linear("at.pose1", f(m.pose_encoder[0].weight), f(m.pose_encoder[0].bias), True, dev)
        self.pose2 = Linear("at.pose2", f(m.pose_encoder[2].weight), f(m.pose_encoder[2].bias), True, dev)
        # LSTM (util.py:557): layer 0's input is [image 512 | audio 256 | pose 256]; the image part is constant
        # over the clip, so its projection (plus both biases) is computed once per sequence and row-broadcast
        self.lstm = []
        for l in range(3):
            wi, wh = f(getattr(m.lstm, "weight_ih_l%d" % l)), f(getattr(m.lstm, "weight_hh_l%d" % l))
            b = f(getattr(m.lstm, "bias_ih_l%d" % l)) + f(getattr(m.lstm, "bias_hh_l%d" % l))
            if l == 0:
                proj = (Linear("at.lstm0.img", wi[:, :512], b, False, dev), Linear("at.lstm0.x", wi[:, 512:], None, False, dev))
            else:
                proj = (None, Linear("at.lstm%d.x" % l, wi, b, False, dev))
            self.lstm.append((proj, wh.contiguous()))
        # decon (util.py:559-575)
        d = m.decon
        s, t = bn_affine(_bn(d[1]))
        w0 = f(d[0].weight)[:, :, 1:5, 1:5] * s.view(1, -1, 1, 1)                 # 1x1 input: out(y,x) uses tap (y+1,x+1)
        b0 = f(d[0].bias) * s + t
        # decon stack on tensor cores: the 1x1 -> 4x4 layer then writes NCHW (columns (c, y, x)) and eamm_nchw_to_act
        # splits it into the bf16 hi/lo planes the convolutions read
        self.dec_tc = os.environ.get("EAMM_B200_AT_DECON", "tc") != "simt"
        if self.dec_tc:
            self.dec0 = Linear("at.decon0", w0.permute(1, 2, 3, 0).reshape(256 * 16, 256), b0.repeat_interleave(16), True, dev)
        else:
            self.dec0 = Linear("at.decon0", w0.permute(2, 3, 1, 0).reshape(16 * 256, 256), b0.repeat(16), True, dev)
        self.dec = []
        for i in (3, 6, 9, 12):
            w, b = f(d[i].weight), f(d[i].bias)                                   # [cin][cout][4][4]
            if i != 12:
                s, t = bn_affine(_bn(d[i + 1]))
                w, b = w * s.view(1, -1, 1, 1), b * s + t
            lay = ConvLayer("at.decon%d" % i, L.CONV_UP2_3X3, L.EPI_RELU if i != 12 else 0,
                            w.permute(1, 0, 2, 3), b, w.shape[0], 16 if self.dec_tc else 4, "tc3" if self.dec_tc else "simt",
                            parity=convT_parity_weights(w))
            lay.flops_per_in_pixel = 2.0 * w.shape[0] * w.shape[1] * 16
            self.dec.append(lay)
        self.ws = WorkspaceCache()

    def workspace(self, B, T, H, W):
        key = (B, T, H, W)
        ws = self.ws.get(key)
        if ws is not None:
            return ws
        dev, M = self.device, B * T
        ws = type("WS", (), {})()
        ws.img = [ActBuf(B, H, W, 4, "f32", dev)]
        h, w = H, W
        for lay in self.img:
            h, w = h // 2, w // 2
            ws.img.append(ActBuf(B, h, w, lay.cout, "f32", dev))
        ws.a_in = ActBuf(M, 28, 12, 4, "f32", dev)
        ws.a0 = ActBuf(M, 28, 12, 64, "f32", dev)
        ws.a1 = ActBuf(M, 28, 12, 128, "f32", dev)
        ws.p1 = ActBuf(M, 26, 5, 128, "f32", dev)
        ws.a3 = ActBuf(M, 26, 5, 256, "f32", dev)
        ws.a4 = ActBuf(M, 26, 5, 256, "f32", dev)
        ws.a5 = ActBuf(M, 26, 5, 512, "f32", dev)
        ws.p2 = ActBuf(M, 12, 2, 512, "f32", dev)
        if self.aud_tc:
            # padded twins of the conv operands (zero-initialised; the padding is kept zero by eamm_act_copy)
            pb = lambda h, w, c: ActBuf(M, h, w, c, "bf16x2", dev)
            ws.a0p, ws.a1p = pb(32, 16, 64), pb(32, 16, 128)
            ws.p1p, ws.a3p, ws.a4p, ws.a5p = pb(32, 8, 128), pb(32, 8, 256), pb(32, 8, 256), pb(32, 8, 512)
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        ws.f1, ws.pose_h, ws.x2 = e(M, 2048), e(M, 128), e(M, 512)
        ws.img_proj, ws.gates = e(B, 4 * self.HID), e(M, 4 * self.HID)
        ws.h = [e(M, self.HID) for _ in range(3)]
        ws.d = [ActBuf(M, 4 << i, 4 << i, c, "bf16x2" if self.dec_tc else "f32", dev) for i, c in enumerate((256, 128, 128, 128))]
        ws.d0_nchw = e(M, 256 * 16) if self.dec_tc else None
        self.ws[key] = ws
        return ws

    def run(self, example_image, audio, pose, weight):
        m, lib, dev = self.m, self.lib, self.device
        st = current_stream_ptr()
        B, T = audio.shape[:2]
        M = B * T
        H, W = example_image.shape[2:]
        ws = self.workspace(B, T, H, W)
        # ---- image feature (util.py:583-587)
        a = ws.img[0].act()
        _launch("at.image_to_act", lambda: L.check(lib.eamm_nchw_to_act(example_image.data_ptr(), B, 3, H, W, C.byref(a), st),
                                                   "nchw_to_act"), nbytes=B * 3 * H * W * 4)
        for i, lay in enumerate(self.img):
            lay.launch(lib, st, ws.img[i].act(), out=ws.img[i + 1].act())
        # ---- MFCC encoder over all B*T windows (util.py:589-592)
        a = ws.a_in.act()
        _launch("at.mfcc_to_act", lambda: L.check(lib.eamm_nchw_to_act(audio.data_ptr(), M, 1, 28, 12, C.byref(a), st),
                                                  "nchw_to_act"), nbytes=M * 28 * 12 * 4)
        self.aud[0].launch(lib, st, ws.a_in.act(), out=ws.a0.act())
        if self.aud_tc:
            self._copy(ws.a0, ws.a0p, 28, 12)                       # exact fp32 -> padded bf16 hi/lo
            self.aud[1].launch(lib, st, ws.a0p.act(), out=ws.a1p.act())
            self._copy(ws.a1p, ws.a1, 28, 12)                       # valid region -> exact fp32 for the max-pool
            self._pool(ws.a1, ws.p1, 1, 2)
            self._copy(ws.p1, ws.p1p, 26, 5)
            self.aud[3].launch(lib, st, ws.p1p.act(), out=ws.a3p.act())
            self._copy(ws.a3p, ws.a3p, 26, 5, zero_rest=True)       # the conv wrote its padding: zero it again
            self.aud[4].launch(lib, st, ws.a3p.act(), out=ws.a4p.act())
            self._copy(ws.a4p, ws.a4p, 26, 5, zero_rest=True)
            self.aud[5].launch(lib, st, ws.a4p.act(), out=ws.a5p.act())
            self._copy(ws.a5p, ws.a5, 26, 5)
        else:
            self.aud[1].launch(lib, st, ws.a0.act(), out=ws.a1.act())
            self._pool(ws.a1, ws.p1, 1, 2)
            self.aud[3].launch(lib, st, ws.p1.act(), out=ws.a3.act())
            self.aud[4].launch(lib, st, ws.a3.act(), out=ws.a4.act())
            self.aud[5].launch(lib, st, ws.a4.act(), out=ws.a5.act())
        self._pool(ws.a5, ws.p2, 2, 2)
        # ---- audio FC (x weight) and pose MLP write the two halves of the LSTM input (util.py:592-594)
        self.fc1.launch(lib, st, ws.p2.t.data_ptr(), self.fc1.K, ws.f1.data_ptr(), 2048, M)
        self.fc2.launch(lib, st, ws.f1.data_ptr(), 2048, ws.x2.data_ptr(), 512, M, scale=weight)
        self.pose1.launch(lib, st, pose.data_ptr(), 6, ws.pose_h.data_ptr(), 128, M)
        self.pose2.launch(lib, st, ws.pose_h.data_ptr(), 128, ws.x2.data_ptr() + 256 * 4, 512, M)
        # ---- LSTM (util.py:596-597)
        x, ldx = ws.x2, 512
        for l, ((p_img, p_x), whh) in enumerate(self.lstm):
            add = None
            if p_img is not None:
                p_img.launch(lib, st, ws.img[-1].t.data_ptr(), 512, ws.img_proj.data_ptr(), 4 * self.HID, B)
                add = ws.img_proj
            p_x.launch(lib, st, x.data_ptr(), ldx, ws.gates.data_ptr(), 4 * self.HID, M, add=add, period=T)
            hout = ws.h[l]
            _launch("at.lstm%d" % l, lambda: L.check(lib.eamm_lstm_layer(ws.gates.data_ptr(), whh.data_ptr(), hout.data_ptr(),
                                                                         B, T, self.HID, st), "lstm_layer"),
                    flops=2.0 * M * 4 * self.HID * self.HID)
            x, ldx = hout, self.HID
        # ---- decon (util.py:600-606)
        if self.dec_tc:
            self.dec0.launch(lib, st, x.data_ptr(), self.HID, ws.d0_nchw.data_ptr(), 16 * 256, M)
            a0 = ws.d[0].act()
            _launch("at.decon0_to_act", lambda: L.check(lib.eamm_nchw_to_act(ws.d0_nchw.data_ptr(), M, 256, 4, 4, C.byref(a0), st),
                                                        "nchw_to_act"), nbytes=M * 4096 * 8)
        else:
            self.dec0.launch(lib, st, x.data_ptr(), self.HID, ws.d[0].t.data_ptr(), 16 * 256, M)
        out = torch.empty(B, T, 35, 64, 64, dtype=torch.float32, device=dev)
        for i, lay in enumerate(self.dec):
            if i < 3:
                lay.launch(lib, st, ws.d[i].act(), out=ws.d[i + 1].act())
            else:
                lay.launch(lib, st, ws.d[i].act(), out_nchw=out, out_nchw_c=35)
        return out

    def _copy(self, src, dst, h, w, zero_rest=False):
        a, b = src.act(), dst.act()
        _launch("at.act_copy", lambda: L.check(self.lib.eamm_act_copy(C.byref(a), C.byref(b), h, w, 1 if zero_rest else 0,
                                                                     current_stream_ptr()), "act_copy"),
                nbytes=src.n * h * w * src.c_buf * 8)

    def _pool(self, src, dst, sy, sx):
        a, b = src.act(), dst.act()
        _launch("at.maxpool", lambda: L.check(self.lib.eamm_maxpool(C.byref(a), C.byref(b), 3, sy, sx, current_stream_ptr()),
                                              "maxpool"), nbytes=src.t.numel() * 4)
