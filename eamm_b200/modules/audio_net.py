"""Drop-in for AT_net2 of /root/reference/modules/util.py:514-613: the per-clip
network that turns MFCC windows, head poses and the example image into the [B, T, 35, 64, 64] feature maps that
KPDetector_a reads frame by frame (demo.py:345, :219).

Same constructor (no arguments), parameter names (`down_blocks.*`, `pose_encoder.*`, `audio_eocder.*`,
`audio_eocder_fc.*`, `lstm.*`, `decon.*`) and forward signature.  The reference also owns an unused StyleGAN2
`generator` (only reached with jaco_net == 'gan'; every shipped config sets 'cnn', config/*.yaml `jaco_net`):
its `generator.*` checkpoint entries are accepted and dropped by load_state_dict, and 'gan' raises.
"""
import torch
from torch import nn

from .util import DownBlock2d, _Fused
from .dense_motion import _EngineMixin
from .. import at_engine as _at_engine


def _conv2d(cin, cout):
    """util.py:1740-1753 `conv2d`: Conv2d(bias=False) + BatchNorm2d + ReLU."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class AT_net2(_EngineMixin, nn.Module):
    _engine_cls = _at_engine.ATNet2Engine

    def __init__(self):
        super().__init__()
        self.down_blocks = nn.ModuleList([DownBlock2d(3 if i == 0 else 2 * (2 ** i), 2 * (2 ** (i + 1)),
                                                      kernel_size=3, padding=1) for i in range(8)])
        self.pose_encoder = nn.Sequential(nn.Linear(6, 128), nn.ReLU(True), nn.Linear(128, 256), nn.ReLU(True))
        self.audio_eocder = nn.Sequential(
            _conv2d(1, 64), _conv2d(64, 128), nn.MaxPool2d(3, stride=(1, 2)),
            _conv2d(128, 256), _conv2d(256, 256), _conv2d(256, 512), nn.MaxPool2d(3, stride=(2, 2)))
        self.audio_eocder_fc = nn.Sequential(nn.Linear(1024 * 12, 2048), nn.ReLU(True), nn.Linear(2048, 256), nn.ReLU(True))
        self.lstm = nn.LSTM(256 * 4, 256, 3, batch_first=True)
        self.decon = nn.Sequential(
            nn.ConvTranspose2d(256, 256, kernel_size=6, stride=2, padding=1, bias=True), nn.BatchNorm2d(256), nn.ReLU(True),
            nn.ConvTranspose2d(256, 128, kernel_size=4, stride=2, padding=1, bias=True), nn.BatchNorm2d(128), nn.ReLU(True),
            nn.ConvTranspose2d(128, 128, kernel_size=4, stride=2, padding=1, bias=True), nn.BatchNorm2d(128), nn.ReLU(True),
            nn.ConvTranspose2d(128, 128, kernel_size=4, stride=2, padding=1, bias=True), nn.BatchNorm2d(128), nn.ReLU(True),
            nn.ConvTranspose2d(128, 32 + 3, kernel_size=4, stride=2, padding=1, bias=True))
        self._init_engine_state()
        self._register_load_state_dict_pre_hook(self._drop_stylegan)

    @staticmethod
    def _drop_stylegan(state_dict, prefix, *args):
        for k in [k for k in state_dict if k.startswith(prefix + "generator.")]:
            del state_dict[k]

    def forward(self, example_image, audio, pose, jaco_net, weight):
        if jaco_net != "cnn":
            if jaco_net == "gan":
                raise NotImplementedError("eamm_b200: AT_net2 with jaco_net == 'gan' (StyleGAN2 decoder) is out of scope; "
                                          "every config of the reference uses 'cnn'")
            raise Exception("jaco_net type wrong")                                   # util.py:611
        eng = self._engine(self.lstm.weight_hh_l0)
        dev = self.lstm.weight_hh_l0.device
        for name, t, shape in (("example_image", example_image, (None, 3, None, None)), ("audio", audio, (None, None, 28, 12)),
                               ("pose", pose, (None, None, 6))):
            if t.device != dev or t.dtype != torch.float32 or t.dim() != len(shape) or \
                    any(s is not None and s != d for s, d in zip(shape, t.shape)):
                raise RuntimeError("eamm_b200: AT_net2 %s must be an fp32 %s tensor on the module's device" % (name, list(shape)))
        B, T = audio.shape[:2]
        if example_image.shape[0] != B or pose.shape[:2] != (B, T):
            raise RuntimeError("eamm_b200: AT_net2 batch/clip sizes disagree")
        if example_image.shape[2] != 256 or example_image.shape[3] != 256:
            raise RuntimeError("eamm_b200: AT_net2 needs a 256x256 example image (its 8 DownBlocks end at 1x1, util.py:587)")
        with torch.no_grad(), torch.cuda.device(dev):
            if T == 0 or B == 0:
                return torch.empty(B, T, 35, 64, 64, device=dev)
            return eng.run(example_image.contiguous(), audio.contiguous(), pose.contiguous(), float(weight))
