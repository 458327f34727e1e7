"""eamm_b200 -- B200-native (sm_100a) implementation of EAMM's per-frame generation hot path.

Public surface mirrors the reference's operator API for this path:
    eamm_b200.modules.generator.OcclusionAwareGenerator      (reference modules/generator.py:8)
    eamm_b200.modules.dense_motion.DenseMotionNetwork        (reference modules/dense_motion.py:7)
Both are nn.Modules with the reference's constructor, state_dict layout and
``forward(source_image, kp_driving, kp_source)``; the work is done by hand-written CUDA kernels in
``eamm_b200/csrc`` behind the C ABI of ``include/eamm_b200.h``.  There is no CPU fallback.
"""
from .config import get_config, FULL_CONFIG, TINY_CONFIG  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("OcclusionAwareGenerator", "DenseMotionNetwork"):
        from .modules.generator import OcclusionAwareGenerator
        from .modules.dense_motion import DenseMotionNetwork
        return {"OcclusionAwareGenerator": OcclusionAwareGenerator, "DenseMotionNetwork": DenseMotionNetwork}[name]
    raise AttributeError(name)
