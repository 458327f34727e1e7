"""Constructor kwargs of the hot path, as the reference's YAML gives them.

Values restate /root/reference/config/MEAD_emo_video_aug_delta_4_crop_random_crop.yaml:26-52
(generator_params + common_params; identical in all four shipped configs).  ``TINY_CONFIG`` is a
shrunken instance of the same architecture used by fast parity tests and the committed golden
fixtures (full tensors fit in a few hundred KB).
"""
import copy

FULL_CONFIG = {
    "num_channels": 3,
    "num_kp": 10,
    "estimate_jacobian": True,
    "block_expansion": 64,
    "max_features": 512,
    "num_down_blocks": 2,
    "num_bottleneck_blocks": 6,
    "estimate_occlusion_map": True,
    "dense_motion_params": {
        "block_expansion": 64,
        "max_features": 1024,
        "num_blocks": 5,
        "scale_factor": 0.25,
    },
}

# Same topology, fewer channels / blocks; image size is chosen by the caller (64x64 in tests).
TINY_CONFIG = {
    "num_channels": 3,
    "num_kp": 3,
    "estimate_jacobian": True,
    "block_expansion": 16,
    "max_features": 64,
    "num_down_blocks": 2,
    "num_bottleneck_blocks": 2,
    "estimate_occlusion_map": True,
    "dense_motion_params": {
        "block_expansion": 16,
        "max_features": 64,
        "num_blocks": 3,
        "scale_factor": 0.25,
    },
}


# KPDetector / KPDetector_a constructor kwargs: kp_detector_params + common_params / audio_params of
# /root/reference/config/MEAD_emo_video_aug_delta_4_crop_random_crop.yaml:27-41 (demo.py:59-66).
FULL_KP_CONFIG = {
    "temperature": 0.1, "block_expansion": 32, "max_features": 1024, "scale_factor": 0.25, "num_blocks": 5,
    "num_kp": 10, "num_channels": 3, "estimate_jacobian": True,
}
TINY_KP_CONFIG = {
    "temperature": 0.1, "block_expansion": 8, "max_features": 32, "scale_factor": 0.25, "num_blocks": 2,
    "num_kp": 3, "num_channels": 3, "estimate_jacobian": True,
}


def get_kp_config(name="full", audio=False):
    """kwargs of KPDetector (audio=False) or KPDetector_a (audio=True: adds num_channels_a, yaml:31-35)."""
    cfg = copy.deepcopy({"full": FULL_KP_CONFIG, "tiny": TINY_KP_CONFIG}[name])
    if audio:
        cfg["num_channels_a"] = 3
    return cfg


def get_config(name="full"):
    """"full" / "tiny", or a constructor corner of the tiny config no shipped YAML uses but the reference's code
    supports: "tiny_sf05" (motion grid 2x the encoder feature grid: generator.py:53-56, :82-83 resize flow and
    occlusion), "tiny_sf1" (dense_motion.py:82: no anti-alias module), "tiny_nodm" (generator.py:20-24,67:
    dense_motion_params=None)."""
    if name in ("full", "tiny"):
        return copy.deepcopy({"full": FULL_CONFIG, "tiny": TINY_CONFIG}[name])
    cfg = copy.deepcopy(TINY_CONFIG)
    if name == "tiny_sf05":
        cfg["dense_motion_params"]["scale_factor"] = 0.5
    elif name == "tiny_sf1":
        cfg["dense_motion_params"]["scale_factor"] = 1
    elif name == "tiny_nodm":
        cfg["dense_motion_params"] = None
        cfg["estimate_occlusion_map"] = False
    else:
        raise KeyError(name)
    return cfg
