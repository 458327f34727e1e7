"""Parameter-holding building blocks with the reference's names and constructor signatures.

These mirror /root/reference/modules/util.py:858-1052 so that ``state_dict()`` /
``load_state_dict()`` / ``print(module)`` behave exactly like the reference's generator
(196 keys, e.g. ``bottleneck.r0.norm1.running_mean``).  They are containers only: the arithmetic
of every block is executed by the fused sm_100a kernels driven from
``eamm_b200.engine`` (BatchNorm is folded into the convolutions there), never by these classes.
"""
import torch
from torch import nn

from ..synth import aa_kernel

BatchNorm2d = nn.BatchNorm2d   # the reference's SynchronizedBatchNorm2d is a _BatchNorm subclass
                               # with identical parameters/buffers (sync_batchnorm/batchnorm.py:38-46)


class _Fused(nn.Module):
    def forward(self, *args, **kwargs):
        raise RuntimeError(
            "%s is a parameter container in eamm_b200; it only runs fused inside "
            "OcclusionAwareGenerator / DenseMotionNetwork.forward" % type(self).__name__)


class ResBlock2d(_Fused):
    """util.py:858-880."""

    def __init__(self, in_features, kernel_size, padding):
        super().__init__()
        self.conv1 = nn.Conv2d(in_features, in_features, kernel_size=kernel_size, padding=padding)
        self.conv2 = nn.Conv2d(in_features, in_features, kernel_size=kernel_size, padding=padding)
        self.norm1 = BatchNorm2d(in_features, affine=True)
        self.norm2 = BatchNorm2d(in_features, affine=True)


class UpBlock2d(_Fused):
    """util.py:883-900."""

    def __init__(self, in_features, out_features, kernel_size=3, padding=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(in_features, out_features, kernel_size=kernel_size, padding=padding, groups=groups)
        self.norm = BatchNorm2d(out_features, affine=True)


class DownBlock2d(_Fused):
    """util.py:903-920."""

    def __init__(self, in_features, out_features, kernel_size=3, padding=1, groups=1):
        super().__init__()
        self.conv = nn.Conv2d(in_features, out_features, kernel_size=kernel_size, padding=padding, groups=groups)
        self.norm = BatchNorm2d(out_features, affine=True)
        self.pool = nn.AvgPool2d(kernel_size=(2, 2))


class SameBlock2d(_Fused):
    """util.py:923-938."""

    def __init__(self, in_features, out_features, groups=1, kernel_size=3, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(in_features, out_features, kernel_size=kernel_size, padding=padding, groups=groups)
        self.norm = BatchNorm2d(out_features, affine=True)


class Encoder(_Fused):
    """util.py:941-960."""

    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        down_blocks = []
        for i in range(num_blocks):
            down_blocks.append(DownBlock2d(in_features if i == 0 else min(max_features, block_expansion * (2 ** i)),
                                           min(max_features, block_expansion * (2 ** (i + 1))),
                                           kernel_size=3, padding=1))
        self.down_blocks = nn.ModuleList(down_blocks)


class Decoder(_Fused):
    """util.py:963-987."""

    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        up_blocks = []
        for i in range(num_blocks)[::-1]:
            in_filters = (1 if i == num_blocks - 1 else 2) * min(max_features, block_expansion * (2 ** (i + 1)))
            out_filters = min(max_features, block_expansion * (2 ** i))
            up_blocks.append(UpBlock2d(in_filters, out_filters, kernel_size=3, padding=1))
        self.up_blocks = nn.ModuleList(up_blocks)
        self.out_filters = block_expansion + in_features


class Hourglass(_Fused):
    """util.py:990-1002."""

    def __init__(self, block_expansion, in_features, num_blocks=3, max_features=256):
        super().__init__()
        self.encoder = Encoder(block_expansion, in_features, num_blocks, max_features)
        self.decoder = Decoder(block_expansion, in_features, num_blocks, max_features)
        self.out_filters = self.decoder.out_filters


class AntiAliasInterpolation2d(_Fused):
    """util.py:1005-1052: fixed 13x13 Gaussian (sigma hard-coded to 1.5), buffer name ``weight``."""

    def __init__(self, channels, scale):
        super().__init__()
        self.register_buffer("weight", aa_kernel(channels, sigma=1.5))
        self.groups = channels
        self.scale = scale
        self.int_inv_scale = int(1 / scale)


def __getattr__(name):
    """`modules.util.AT_net2` (util.py:514) lives in audio_net.py; resolved lazily to keep the import graph acyclic."""
    if name == "AT_net2":
        from .audio_net import AT_net2
        return AT_net2
    raise AttributeError(name)
