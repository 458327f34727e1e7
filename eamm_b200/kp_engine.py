"""Executor of the keypoint-detector heads (reference
/root/reference/modules/keypoint_detector.py:77-105 and :180-205)."""
import ctypes as C

import torch

from . import _lib as L
from .engine import (ActBuf, ConvLayer, HourglassPlan, WorkspaceCache, _launch, _round_up, current_stream_ptr, impl_for)


class KPDetectorEngine:
    def __init__(self, module, precision):
        m = self.m = module
        self.lib = L.load()
        self.impl, self.mode, self.calign, self.nalign = impl_for(precision)
        dev = self.device = m.kp.weight.device
        ca = self.calign
        feat = m.kp.in_channels
        wk, bk = m.kp.weight.detach().float(), m.kp.bias.detach().float()
        if m.jacobian is not None:            # kp and jacobian 7x7 convs merged into one (K + 4J couts)
            wk = torch.cat([wk, m.jacobian.weight.detach().float()], 0)
            bk = torch.cat([bk, m.jacobian.bias.detach().float()], 0)
        self.hg = None
        if m.uses_predictor:
            cin0 = m.predictor.encoder.down_blocks[0].conv.in_channels
            if cin0 != 3:
                raise RuntimeError("eamm_b200: the B200 keypoint detector supports num_channels == 3 only")
            self.hg = HourglassPlan(m.predictor, cin0, ca, self.nalign, self.impl, prefix="kp.hg")
            wk, cin_slot = self.hg.split_cat_weights(wk, self.hg.dec_ch[-1], cin0)
            self.step = int(1 / m.scale_factor) if m.scale_factor != 1 else 1
            self.taps, self.g1 = 1, torch.ones(1, dtype=torch.float32, device=dev)
            if self.step != 1:
                k2 = m.down.weight.detach().float()[0, 0]
                if k2.shape[0] > 13 or k2.shape[0] % 2 == 0:
                    raise RuntimeError("eamm_b200: anti-alias kernel must be odd and at most 13x13")
                g1 = k2.sum(1)
                self.g1 = (g1 / g1.sum()).contiguous()
                self.taps = k2.shape[0]
        else:
            cin_slot = _round_up(feat, ca)
        self.head = ConvLayer("kp_head", L.CONV_7X7, 0, wk, bk, cin_slot, self.nalign, self.impl, cin_valid=feat)
        self.ws = WorkspaceCache()

    def workspace(self, B, h, w):
        ws = self.ws.get((B, h, w))
        if ws is None:
            ws = type("WS", (), {})()
            if self.hg is not None:
                ws.cat, ws.bott = self.hg.buffers(B, h, w, self.mode, self.device)
            else:
                ws.feat = ActBuf(B, h, w, self.head.cin, self.mode, self.device)
            ws.logits = torch.empty(B, h, w, self.head.cout, dtype=torch.float32, device=self.device)
            self.ws[(B, h, w)] = ws
        return ws

    def run(self, x):
        m, lib, dev = self.m, self.lib, self.device
        st = current_stream_ptr()
        B, Cc, H, W = x.shape
        K, J = m.num_kp, m.num_jacobian_maps
        if self.hg is not None:
            h, w = H // self.step, W // self.step
            ws = self.workspace(B, h, w)
            cat0 = ws.cat[0]
            dst = cat0.act(c_off=cat0.s_up, c=cat0.s_sk)
            _launch("kp.aa_downsample", lambda: L.check(
                lib.eamm_aa_downsample_act(x.data_ptr(), Cc * H * W, B, H, W, self.step, self.g1.data_ptr(), self.taps,
                                           C.byref(dst), st), "aa_downsample_act"), nbytes=B * Cc * H * W * 4)
            self.hg.run(lib, st, ws.cat, ws.bott)
            feat = cat0.act()
        else:
            h, w = H, W
            if Cc != m.kp.in_channels:
                raise RuntimeError("eamm_b200: feature map must have %d channels" % m.kp.in_channels)
            ws = self.workspace(B, h, w)
            fa = ws.feat.act()
            _launch("kp.nchw_to_act", lambda: L.check(
                lib.eamm_nchw_to_act(x.data_ptr(), B, Cc, h, w, C.byref(fa), st), "nchw_to_act"),
                nbytes=B * Cc * h * w * 4)
            feat = ws.feat.act()
        self.head.launch(lib, st, feat, out_nhwc_f32=ws.logits)
        off = 3 - m.pad
        hh, ww = h - 2 * off, w - 2 * off
        heatmap = torch.empty(B, K, hh, ww, dtype=torch.float32, device=dev)
        value = torch.empty(B, K, 2, dtype=torch.float32, device=dev)
        jac = torch.empty(B, K, 2, 2, dtype=torch.float32, device=dev) if J else None
        _launch("kp_head", lambda: L.check(
            lib.eamm_kp_head(ws.logits.data_ptr(), self.head.cout, B, h, w, K, J, m.pad, float(m.temperature),
                             heatmap.data_ptr(), value.data_ptr(), jac.data_ptr() if J else None, st), "kp_head"),
            nbytes=B * h * w * self.head.cout * 4 * 3)
        out = {"value": value, "heatmap": heatmap}
        if J:
            out["jacobian"] = jac
        return out
