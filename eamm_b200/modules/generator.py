"""Drop-in for /root/reference/modules/generator.py: same constructor, parameters and forward."""
import torch
from torch import nn

from .util import ResBlock2d, SameBlock2d, UpBlock2d, DownBlock2d
from .dense_motion import DenseMotionNetwork, _EngineMixin, check_status
from .. import engine as _engine


class OcclusionAwareGenerator(_EngineMixin, nn.Module):
    """
    Generator that given source image and and keypoints try to transform image according to movement trajectories
    induced by keypoints. Generator follows Johnson architecture (reference: modules/generator.py:8-97).
    """
    _engine_cls = _engine.GeneratorEngine

    def __init__(self, num_channels, num_kp, block_expansion, max_features, num_down_blocks,
                 num_bottleneck_blocks, estimate_occlusion_map=False, dense_motion_params=None, estimate_jacobian=False):
        super().__init__()
        if dense_motion_params is not None:
            self.dense_motion_network = DenseMotionNetwork(num_kp=num_kp, num_channels=num_channels,
                                                           estimate_occlusion_map=estimate_occlusion_map,
                                                           **dense_motion_params)
        else:
            self.dense_motion_network = None
        self.first = SameBlock2d(num_channels, block_expansion, kernel_size=(7, 7), padding=(3, 3))
        down_blocks = []
        for i in range(num_down_blocks):
            in_features = min(max_features, block_expansion * (2 ** i))
            out_features = min(max_features, block_expansion * (2 ** (i + 1)))
            down_blocks.append(DownBlock2d(in_features, out_features, kernel_size=(3, 3), padding=(1, 1)))
        self.down_blocks = nn.ModuleList(down_blocks)
        up_blocks = []
        for i in range(num_down_blocks):
            in_features = min(max_features, block_expansion * (2 ** (num_down_blocks - i)))
            out_features = min(max_features, block_expansion * (2 ** (num_down_blocks - i - 1)))
            up_blocks.append(UpBlock2d(in_features, out_features, kernel_size=(3, 3), padding=(1, 1)))
        self.up_blocks = nn.ModuleList(up_blocks)
        self.bottleneck = torch.nn.Sequential()
        in_features = min(max_features, block_expansion * (2 ** num_down_blocks))
        for i in range(num_bottleneck_blocks):
            self.bottleneck.add_module('r' + str(i), ResBlock2d(in_features, kernel_size=(3, 3), padding=(1, 1)))
        self.final = nn.Conv2d(block_expansion, num_channels, kernel_size=(7, 7), padding=(3, 3))
        self.estimate_occlusion_map = estimate_occlusion_map
        self.num_channels = num_channels
        # opt-in: reuse the encoder feature maps while the caller keeps passing the same source tensor
        # (same storage, same version counter) -- what demo.py:279 does for a whole clip
        self.cache_source = False
        # opt-in: also return 'prediction_u8' [B,H,W,C] uint8 = img_as_ubyte(prediction), the
        # frames demo.py:281,507 assembles on the host; written by the final conv's epilogue
        self.emit_u8 = False
        self._init_engine_state()

    def forward(self, source_image, kp_driving, kp_source):
        eng = self._engine(self.final.weight)
        if source_image.device != self.final.weight.device or source_image.dtype != torch.float32:
            raise RuntimeError("eamm_b200: source_image must be an fp32 tensor on the module's device")
        with torch.no_grad(), torch.cuda.device(source_image.device):
            out = eng.run(source_image, kp_driving, kp_source)
            if self.strict_errors and eng.dm is not None:
                check_status(eng.dm.last_status)
            if self.strict_errors and eng.needs_recalibration():
                # an activation maximum left the window its fp16 + fp8 pre-scale was calibrated for: the exponents
                # have been re-centred on this batch, run it again (rare: the window spans 4 octaves)
                out = eng.run(source_image, kp_driving, kp_source)
        return out
