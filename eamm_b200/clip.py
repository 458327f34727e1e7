"""Per-clip keypoint glue on the device.

`smooth_and_normalize` is the batched equivalent of demo.py:228-278: One-Euro smoothing of the per-frame
detector outputs, the emotion-row accumulation and `normalize_kp`, for all T frames in one kernel.  Its
result feeds `OcclusionAwareGenerator.forward` as a batch of T driving keypoints.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .engine import current_stream_ptr

EMO_ROWS = ((1, 0, 0.2), (4, 1, 1.0), (6, 2, 1.0))       # demo.py:266-271 ('linear_3')
KP_FILTER = (0.05, 8.0, 1.0, 100.0, 10.0)                 # mincutoff, beta, dcutoff, freq, scale (demo.py:241-245)
EMO_FILTER = (1.0, 0.2, 1.0, 100.0, 100.0)                # demo.py:231-236


def mfcc_windows(mfcc):
    """The audio windows `test_auido` cuts from a whole utterance (demo.py:318-333): `mfcc` [n,13] is
    python_speech_features.mfcc(speech, 16000, winstep=0.01) of the zero-padded waveform; frame `ind` (3 <= ind <=
    n // 4 - 4) sees rows (ind-3)*4 .. (ind+4)*4 without cepstral coefficient 0.  Returns [T,28,12] fp32, T = n // 4 - 6:
    one strided view instead of the reference's Python loop, same rows."""
    mfcc = np.ascontiguousarray(np.asarray(mfcc, dtype=np.float32))
    if mfcc.ndim != 2 or mfcc.shape[1] < 2:
        raise ValueError("eamm_b200: mfcc must be [n, n_cep >= 2]")
    T = mfcc.shape[0] // 4 - 6
    if T <= 0:
        return np.zeros((0, 28, mfcc.shape[1] - 1), dtype=np.float32)
    s0, s1 = mfcc.strides
    win = np.lib.stride_tricks.as_strided(mfcc, shape=(T, 28, mfcc.shape[1]), strides=(4 * s0, s0, s1), writeable=False)
    return np.ascontiguousarray(win[:, :, 1:])


def clip_inputs_from_windows(mfcc13, pose7, T=None):
    """Host-side input preparation of `test_auido` (demo.py:286-343) for pre-windowed MFCC such as the LRW samples
    (/root/reference/dataset/LRW/MFCC/*/*.npy, [n,28,13]): drop cepstral coefficient 0 (`[:, :, 1:]`, demo.py:329),
    tile the windows to T frames; pose [m,7] -> first six columns (demo.py:297), mirrored and tiled until it covers
    the clip, then cut to T rows (demo.py:334-340).  Returns (mfcc [1,T,28,12], pose [1,T,6]) fp32 CPU tensors."""
    mfcc13 = np.asarray(mfcc13)
    pose = np.asarray(pose7)[:, :6]
    T = mfcc13.shape[0] if T is None else T
    mfcc = np.tile(mfcc13[:, :, 1:], (-(-T // mfcc13.shape[0]), 1, 1))[:T]
    if len(pose) == 1:
        pose = np.repeat(pose, 100, 0)                                   # demo.py:298-299
    if len(pose) < T:
        gap = T - len(pose)
        n = int((gap / len(pose) / 2)) + 2
        pose = np.tile(np.concatenate((pose, pose[::-1, :]), axis=0), (n, 1))
    pose = pose[:T]
    return (torch.from_numpy(mfcc.astype(np.float32)).unsqueeze(0).contiguous(),
            torch.from_numpy(pose.astype(np.float32)).unsqueeze(0).contiguous())


def movement_scale(kp_source, kp_driving_initial):
    """sqrt(area(source hull)) / sqrt(area(initial driving hull)), demo.py:114-117 (host side, once per clip)."""
    from scipy.spatial import ConvexHull
    sa = ConvexHull(kp_source["value"][0].detach().cpu().numpy()).volume
    da = ConvexHull(kp_driving_initial["value"][0].detach().cpu().numpy()).volume
    return float(np.sqrt(sa) / np.sqrt(da))


def smooth_and_normalize(kp_driving_all, kp_source, kp_driving_initial, emo_driving_all=None, relative=True,
                         scale=1.0, emo_rows=EMO_ROWS):
    """kp_driving_all / emo_driving_all: {'value': [T,K,2], 'jacobian': [T,K,2,2]} CUDA fp32 (the stacked
    per-frame detector outputs); kp_source / kp_driving_initial: batch-1 dicts.  Returns the batch of T
    normalised driving keypoints."""
    dv, dj = kp_driving_all["value"].contiguous(), kp_driving_all["jacobian"].contiguous()
    if dv.device.type != "cuda" or dv.dtype != torch.float32:
        raise RuntimeError("eamm_b200: keypoints must be fp32 CUDA tensors")
    T, K = dv.shape[:2]
    lib = L.load()
    dev = dv.device
    out_v, out_j = torch.empty_like(dv), torch.empty_like(dj)
    ev = ej = scratch = rows = gains = None
    Ke = 0
    if emo_driving_all is not None:
        ev, ej = emo_driving_all["value"].contiguous(), emo_driving_all["jacobian"].contiguous()
        Ke = ev.shape[1]
        scratch = torch.empty(T * Ke * 6, dtype=torch.float32, device=dev)
        rows = torch.tensor([[d, s] for d, s, _ in emo_rows], dtype=torch.int32, device=dev)
        gains = torch.tensor([g for _, _, g in emo_rows], dtype=torch.float32, device=dev)
    fk, fe = L.OneEuro(*KP_FILTER), L.OneEuro(*EMO_FILTER)
    sv, sj = kp_source["value"][:1].contiguous(), kp_source["jacobian"][:1].contiguous()
    iv, ij = kp_driving_initial["value"][:1].contiguous(), kp_driving_initial["jacobian"][:1].contiguous()
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        L.check(lib.eamm_kp_clip(p(dv), p(dj), p(ev), p(ej), T, K, Ke, C.byref(fk), C.byref(fe), p(rows), p(gains),
                                 len(emo_rows) if ev is not None else 0, p(sv), p(sj), p(iv), p(ij), float(scale),
                                 1 if relative else 0, p(out_v), p(out_j), p(scratch), current_stream_ptr()), "kp_clip")
    return {"value": out_v, "jacobian": out_j}


def animate_audio_clip(at_net, kp_detector, kp_detector_a, generator, example_image, mfcc, pose, weight=1.6,
                       relative=True, adapt_movement_scale=True, emo_driving_all=None, chunk=64, emit_u8=True):
    """One audio-driven clip, batched over its T frames: `test_auido` + `make_animation_smooth` of the reference
    (demo.py:345 and :206-281) without the per-frame host round trips.

    example_image [1,3,256,256] fp32 CUDA in [0,1]; mfcc [1,T,28,12] (the windows of demo.py:323-328); pose [1,T,6].
    emo_driving_all: optional stacked outputs of the (out-of-scope) emotion detector, as `smooth_and_normalize` takes.
    Returns the frames as uint8 [T,256,256,3] (emit_u8; what demo.py:507 writes) or fp32 [T,3,256,256].
    """
    T = mfcc.shape[1]
    deco = at_net(example_image, mfcc, pose, "cnn", weight)                      # demo.py:345
    kp_source = kp_detector(example_image)                                       # demo.py:206
    kp_all = kp_detector_a(deco[0])                                              # demo.py:207,219 for every frame at once
    kp_init = {k: kp_all[k][:1] for k in ("value", "jacobian")}
    scale = movement_scale(kp_source, kp_init) if adapt_movement_scale else 1.0  # demo.py:114-119
    kp_norm = smooth_and_normalize(kp_all, kp_source, kp_init, emo_driving_all, relative=relative, scale=scale)
    frames = []
    prev = generator.emit_u8
    generator.emit_u8 = bool(emit_u8)
    try:
        for t0 in range(0, T, chunk):
            n = min(chunk, T - t0)
            out = generator(example_image[:1].expand(n, -1, -1, -1),
                            kp_driving={k: v[t0:t0 + n] for k, v in kp_norm.items()},
                            kp_source={k: kp_source[k][:1].expand(n, *kp_source[k].shape[1:]) for k in ("value", "jacobian")})
            frames.append(out["prediction_u8"] if emit_u8 else out["prediction"])
    finally:
        generator.emit_u8 = prev
    return torch.cat(frames, 0)
