"""Drop-in for /root/reference/modules/dense_motion.py: same constructor, parameters and forward."""
import os

import torch
from torch import nn

from .util import Hourglass, AntiAliasInterpolation2d
from .. import engine as _engine


class _EngineMixin:
    """Lazy (re)build of the packed-weight engine; invalidated whenever parameters may have changed."""

    _engine_cls = None

    def _init_engine_state(self):
        self._eng = None
        self._precision = os.environ.get("EAMM_B200_PRECISION", "fp32")
        self.strict_errors = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        self._eng = None
        for child in self.children():
            if isinstance(child, _EngineMixin):
                child._invalidate()

    def refresh_weights(self):
        """Drop the packed weights (and calibration state); the next forward repacks from the current parameters.
        `load_state_dict`, `.to()` / `.cuda()` and a `precision` change do this by themselves; call it after editing
        parameters or BatchNorm statistics IN PLACE (`p.data.copy_(...)`), which nothing can observe cheaply."""
        self._invalidate()

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, value):
        if value not in _engine.PRECISIONS:
            raise ValueError("precision must be one of %s" % (_engine.PRECISIONS,))
        if value != self._precision:
            self._precision = value
            self._eng = None

    def _apply(self, fn, *args, **kwargs):      # .cuda() / .to() / .float() move the parameters
        self._eng = None
        return super()._apply(fn, *args, **kwargs)

    def _engine(self, ref_param):
        if self.training:
            # (train mode would need batch statistics in every BatchNorm; the kernels fold the running statistics)
            raise RuntimeError("eamm_b200 implements the inference path only: call .eval() and run under "
                               "torch.no_grad() (the reference's demo.py:105,195 does)")
        if ref_param.device.type != "cuda":
            raise RuntimeError("eamm_b200 has no CPU path: move the module to a CUDA (sm_100) device")
        if self._eng is not None and getattr(self._eng, "device", ref_param.device) != ref_param.device:
            self._eng = None            # e.g. an nn.DataParallel replica that inherited the original's engine
        if self._eng is None:
            with torch.no_grad(), torch.cuda.device(ref_param.device):
                self._eng = self._engine_cls(self, self._precision)
        return self._eng


class DenseMotionNetwork(_EngineMixin, nn.Module):
    """
    Module that predicting a dense motion from sparse motion representation given by kp_source and kp_driving
    (reference: modules/dense_motion.py:7-113).
    """
    _engine_cls = _engine.DenseMotionEngine

    def __init__(self, block_expansion, num_blocks, max_features, num_kp, num_channels, estimate_occlusion_map=False,
                 scale_factor=1, kp_variance=0.01):
        super().__init__()
        self.hourglass = Hourglass(block_expansion=block_expansion, in_features=(num_kp + 1) * (num_channels + 1),
                                   max_features=max_features, num_blocks=num_blocks)
        self.mask = nn.Conv2d(self.hourglass.out_filters, num_kp + 1, kernel_size=(7, 7), padding=(3, 3))
        if estimate_occlusion_map:
            self.occlusion = nn.Conv2d(self.hourglass.out_filters, 1, kernel_size=(7, 7), padding=(3, 3))
        else:
            self.occlusion = None
        self.num_kp = num_kp
        self.num_channels = num_channels
        self.scale_factor = scale_factor
        self.kp_variance = kp_variance
        if self.scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, self.scale_factor)
        self._init_engine_state()

    def forward(self, source_image, kp_driving, kp_source):
        eng = self._engine(self.mask.weight)
        if source_image.device != self.mask.weight.device or source_image.dtype != torch.float32:
            raise RuntimeError("eamm_b200: source_image must be an fp32 tensor on the module's device")
        with torch.no_grad(), torch.cuda.device(source_image.device):
            src = source_image if source_image.stride(0) == 0 else source_image.contiguous()
            if src.stride(0) == 0:
                src = src[:1].contiguous().expand(source_image.shape[0], -1, -1, -1)
            out, ws = eng.run(src, kp_driving, kp_source)
            if self.strict_errors:
                check_status(ws.status)
        return out


def check_status(status, clear=False):
    """Raise like torch.inverse does (dense_motion.py:56) if a driving Jacobian was singular.  clear: reset the flag
    after reading it (callers that run non-strict forwards and check once, e.g. FramePipeline.drain)."""
    flag = int(status.item())
    if clear and flag:
        status.zero_()
    if flag & 1:
        raise torch.linalg.LinAlgError("eamm_b200: kp_driving['jacobian'] contains a singular matrix "
                                       "(torch.inverse in the reference raises here)")
