"""Drop-ins for /root/reference/modules/keypoint_detector.py: KPDetector (:7-105) and KPDetector_a (:110-205).

Same constructors, parameter names (`predictor.*`, `kp.*`, `jacobian.*`, `down.weight`) and output
dict ({'value', 'heatmap', 'jacobian'}).  This is SURVEY.md section 8(f) rank 1 -- the step right before the
generation hot path; it reuses the Hourglass convolution kernels and adds one fused
softmax-expectation kernel (`eamm_kp_head`).
"""
import torch
from torch import nn

from .util import Hourglass, AntiAliasInterpolation2d
from .dense_motion import _EngineMixin
from .. import kp_engine as _kp_engine


class _KPBase(_EngineMixin, nn.Module):
    def _build(self, block_expansion, num_kp, in_features, num_channels, max_features, num_blocks, temperature,
               estimate_jacobian, scale_factor, single_jacobian_map, pad):
        self.predictor = Hourglass(block_expansion, in_features=in_features, max_features=max_features,
                                   num_blocks=num_blocks)
        self.kp = nn.Conv2d(in_channels=self.predictor.out_filters, out_channels=num_kp, kernel_size=(7, 7),
                            padding=pad)
        if estimate_jacobian:
            self.num_jacobian_maps = 1 if single_jacobian_map else num_kp
            self.jacobian = nn.Conv2d(in_channels=self.predictor.out_filters,
                                      out_channels=4 * self.num_jacobian_maps, kernel_size=(7, 7), padding=pad)
            self.jacobian.weight.data.zero_()                                        # keypoint_detector.py:27-28
            self.jacobian.bias.data.copy_(torch.tensor([1, 0, 0, 1] * self.num_jacobian_maps, dtype=torch.float))
        else:
            self.jacobian = None
            self.num_jacobian_maps = 0
        self.num_kp = num_kp
        self.pad = pad
        self.temperature = temperature
        self.scale_factor = scale_factor
        if self.scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, self.scale_factor)
        self._init_engine_state()

    def _run(self, x):
        eng = self._engine(self.kp.weight)
        if x.device != self.kp.weight.device or x.dtype != torch.float32 or x.dim() != 4:
            raise RuntimeError("eamm_b200: input must be an fp32 [B,C,H,W] tensor on the module's device")
        with torch.no_grad(), torch.cuda.device(x.device):
            return eng.run(x.contiguous())


class KPDetector(_KPBase):
    """
    Detecting a keypoints. Return keypoint position and jacobian near each keypoint.
    """
    _engine_cls = _kp_engine.KPDetectorEngine

    def __init__(self, block_expansion, num_kp, num_channels, max_features,
                 num_blocks, temperature, estimate_jacobian=False, scale_factor=1,
                 single_jacobian_map=False, pad=0):
        super().__init__()
        self.uses_predictor = True
        self._build(block_expansion, num_kp, num_channels, num_channels, max_features, num_blocks, temperature,
                    estimate_jacobian, scale_factor, single_jacobian_map, pad)

    def forward(self, x):
        return self._run(x)


class KPDetector_a(_KPBase):
    """
    Detecting a keypoints. Return keypoint position and jacobian near each keypoint
    (audio branch: forward() takes the [B, block_expansion + num_channels_a, h, w] feature map directly).
    """
    _engine_cls = _kp_engine.KPDetectorEngine

    def __init__(self, block_expansion, num_kp, num_channels, num_channels_a, max_features,
                 num_blocks, temperature, estimate_jacobian=False, scale_factor=1,
                 single_jacobian_map=False, pad=0):
        super().__init__()
        self.uses_predictor = False
        self._build(block_expansion, num_kp, num_channels_a, num_channels, max_features, num_blocks, temperature,
                    estimate_jacobian, scale_factor, single_jacobian_map, pad)

    def forward(self, feature_map):
        return self._run(feature_map)
