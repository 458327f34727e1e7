"""CPU ORACLE (test infrastructure only) for the per-clip keypoint glue between the keypoint detector and
the generator -- SURVEY.md section 8(f) rank 2.  Restates, for a whole clip of T frames:

  * OneEuroFilter.process / LowPassFilter.process          /root/reference/filter1.py:14-47
  * the smoothing loops over the clip                       /root/reference/demo.py:231-248
      emotion keypoints: mincutoff 1, beta 0.2, freq 100, values scaled by 100
      driving keypoints: mincutoff 0.05, beta 8, freq 100, values scaled by 10
  * the emotion row accumulation                            /root/reference/demo.py:263-271
      kp rows 1, 4, 6 += emotion rows 0 (x0.2), 1, 2   (value and jacobian alike)
  * normalize_kp with relative movement / jacobian          /root/reference/demo.py:112-132

PARITY PIN: tools/make_golden.py executes the reference's own OneEuroFilter / normalize_kp source text
(filter1.py and demo.py cannot be imported: matplotlib, dlib, ... are absent) on the same seeded clip and
asserts this restatement reproduces it bit-for-bit before writing tests/golden/kp_glue_*.npz.
"""
import numpy as np
import torch

EMO_ROWS = ((1, 0, 0.2), (4, 1, 1.0), (6, 2, 1.0))        # (kp row, emotion row, gain), demo.py:266-271
KP_FILTER = dict(mincutoff=0.05, beta=8.0, dcutoff=1.0, freq=100.0, scale=10.0)     # demo.py:241-245
EMO_FILTER = dict(mincutoff=1.0, beta=0.2, dcutoff=1.0, freq=100.0, scale=100.0)    # demo.py:231-236


def one_euro_series(frames, mincutoff, beta, dcutoff, freq, scale):
    """filter1.py:28-47 applied to the list `frames` (fp32 tensors), as demo.py:235/244 calls it:
    process(x * scale) / scale per frame.  Python-float scalars act in fp32 on the tensors, the
    per-element cutoff/alpha arithmetic is float32 (numpy weak-scalar rules)."""
    te = 1.0 / freq

    def alpha(cutoff):
        tau = 1.0 / (2 * np.pi * cutoff)
        return 1.0 / (1.0 + tau / te)

    prev_x = prev_dx_f = prev_x_f = None
    out = []
    for v in frames:
        x = v * scale
        dx = torch.zeros_like(x) if prev_x is None else (x - prev_x) * freq
        a_d = alpha(dcutoff)
        edx = dx if prev_dx_f is None else a_d * dx + (1.0 - a_d) * prev_dx_f
        prev_dx_f = edx
        cutoff = mincutoff + beta * np.abs(edx.numpy())
        a = alpha(cutoff)
        xf = x if prev_x_f is None else torch.from_numpy(np.asarray(a)) * x + torch.from_numpy(np.asarray(1.0 - a)) * prev_x_f
        prev_x, prev_x_f = x, xf
        out.append(xf / scale)
    return out


def normalize_kp(kp_source, kp_driving, kp_driving_initial, movement_scale=1.0, relative=True):
    """demo.py:112-132 with use_relative_movement = use_relative_jacobian = relative; `movement_scale` is the
    ConvexHull area ratio of :114-117 (host-side, once per clip)."""
    new = dict(kp_driving)
    if relative:
        diff = (kp_driving["value"] - kp_driving_initial["value"]) * movement_scale
        new["value"] = diff + kp_source["value"]
        jd = torch.matmul(kp_driving["jacobian"], torch.inverse(kp_driving_initial["jacobian"]))
        new["jacobian"] = torch.matmul(jd, kp_source["jacobian"])
    return new


def clip_glue(drv_value, drv_jac, emo_value, emo_jac, kp_source, kp_initial, movement_scale=1.0, relative=True):
    """Whole-clip restatement of demo.py:228-278.  drv_* [T,K,...], emo_* [T,Ke,...] or None -> [T,K,2], [T,K,2,2]."""
    T = drv_value.shape[0]
    split = lambda t: [t[i:i + 1].clone() for i in range(T)]
    kv = one_euro_series(split(drv_value), **KP_FILTER)
    kj = one_euro_series(split(drv_jac), **KP_FILTER)
    if emo_value is not None:
        ev = one_euro_series(split(emo_value), **EMO_FILTER)
        ej = one_euro_series(split(emo_jac), **EMO_FILTER)
    out_v, out_j = [], []
    for t in range(T):
        v, j = kv[t].clone(), kj[t].clone()
        if emo_value is not None:
            for dst, src, gain in EMO_ROWS:
                v[:, dst] = v[:, dst] + ev[t][:, src] * gain
                j[:, dst] = j[:, dst] + ej[t][:, src] * gain
        n = normalize_kp(kp_source, {"value": v, "jacobian": j}, kp_initial, movement_scale, relative)
        out_v.append(n["value"])
        out_j.append(n["jacobian"])
    return torch.cat(out_v, 0), torch.cat(out_j, 0)
