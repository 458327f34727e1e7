"""The CPU oracle against the golden fixtures generated from the real reference (tools/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from eamm_b200 import get_config, synth
from oracle import eamm_oracle as oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STRIDES = {"mask": 4, "sparse_deformed": 4, "occlusion_map": 4, "deformed": 8, "prediction": 8, "deformation": 4}
CASES = [("tiny_b2", "tiny", True), ("tiny_b3_nojac", "tiny", True),
         # constructor corners (flow/occlusion resize, no anti-alias module, no dense-motion network)
         ("tiny_sf05_b2", "tiny_sf05", True), ("tiny_sf1_b2", "tiny_sf1", True), ("tiny_nodm_b2", "tiny_nodm", True),
         ("full_b2", "full", False),
         ("full_b3_shared", "full", False),
         ("full_b16_shared", "full", False)]       # BASELINE.json configs[0]: one source + 16 kp/jacobian frames


def load_case(name, cfg_name):
    blob = np.load(os.path.join(GOLD, name + ".npz"))
    batch, size, jac, shared = [int(v) for v in blob["meta"]]
    cfg = get_config(cfg_name)
    sd = synth.make_state_dict(cfg, seed=0)
    src, kpd, kps = synth.make_inputs(batch, cfg, size=size, seed=1, with_jacobian=bool(jac), shared_source=bool(shared))
    return blob, cfg, sd, src, kpd, kps


def subsample(k, a, full):
    if full:
        return a
    s = STRIDES[k]
    return a[:, ::s, ::s, :] if k == "deformation" else a[..., ::s, ::s]


@pytest.mark.parametrize("name,cfg_name,full", CASES)
def test_oracle_reproduces_reference_golden(name, cfg_name, full):
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    blob, cfg, sd, src, kpd, kps = load_case(name, cfg_name)
    chk = np.array([src.double().sum(), kpd["value"].double().sum(), kps["value"].double().sum()])
    np.testing.assert_allclose(chk, blob["in_checksum"], rtol=0, atol=0)      # seeded inputs are reproducible
    got = oracle.generator_forward(sd, cfg, src, kpd, kps)
    if cfg.get("dense_motion_params") is not None:
        got["deformation"] = oracle.dense_motion_forward(sd, cfg, src, kpd, kps)["deformation"]
    else:
        assert set(got) == {"prediction"}                                     # generator.py:66-95 without a motion network
    for k, v in got.items():
        a = v.numpy()
        # same torch build on both boxes -> bit-exact; the tolerance only absorbs a different BLAS thread split
        np.testing.assert_allclose(subsample(k, a, full), blob[k], rtol=0, atol=2e-6, err_msg=k)
        s = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
        np.testing.assert_allclose(s, blob["sum_" + k], rtol=1e-6, err_msg="checksum " + k)


NATURAL_STRIDES = {"mask": 2, "sparse_deformed": 2, "occlusion_map": 2, "deformed": 4, "prediction": 4, "deformation": 2}


def natural_case():
    """Natural source images (pixels carried by the fixture, tools/make_golden.py run_natural_case) + seeded keypoints."""
    blob = np.load(os.path.join(GOLD, "natural_b4.npz"))
    cfg = get_config("full")
    src = (torch.from_numpy(blob["pixels_u8"]).float() / 255.0).permute(0, 3, 1, 2).contiguous()
    _, kpd, kps = synth.make_inputs(src.shape[0], cfg, size=256, seed=1)
    return blob, cfg, src, kpd, kps


def natural_subsample(k, a):
    s = NATURAL_STRIDES[k]
    return a[:, ::s, ::s, :] if k == "deformation" else a[..., ::s, ::s]


def test_oracle_reproduces_reference_golden_on_natural_images():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    blob, cfg, src, kpd, kps = natural_case()
    taps = {}
    got = oracle.generator_forward(synth.make_state_dict(cfg, seed=0), cfg, src, kpd, kps, taps=taps)
    got["deformation"] = taps["deformation"]
    for k, v in got.items():
        a = v.numpy()
        np.testing.assert_allclose(natural_subsample(k, a), blob[k], rtol=0, atol=2e-6, err_msg=k)
        s = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
        np.testing.assert_allclose(s, blob["sum_" + k], rtol=1e-6, err_msg="checksum " + k)


def test_structural_known_answers():
    """mask sums to 1, occlusion/prediction in (0,1), identity keypoints give identity flows."""
    cfg = get_config("tiny")
    sd = synth.make_state_dict(cfg, seed=0)
    src, kpd, kps = synth.make_inputs(2, cfg, size=64, seed=3)
    out = oracle.generator_forward(sd, cfg, src, kpd, kps)
    assert torch.allclose(out["mask"].sum(1), torch.ones(2, 16, 16), atol=1e-6)
    assert out["occlusion_map"].min() > 0 and out["occlusion_map"].max() < 1
    assert out["prediction"].min() > 0 and out["prediction"].max() < 1
    same = {"value": kpd["value"], "jacobian": torch.eye(2).expand(2, cfg["num_kp"], 2, 2).contiguous()}
    sm = oracle.sparse_motions(same, same, 16, 16)
    ident = oracle.make_coordinate_grid(16, 16).view(1, 1, 16, 16, 2).expand_as(sm)
    assert torch.allclose(sm, ident, atol=1e-6)


def test_singular_jacobian_raises_like_reference():
    cfg = get_config("tiny")
    src, kpd, kps = synth.make_inputs(1, cfg, size=64, seed=3)
    kpd["jacobian"][0, 0] = 0.0
    with pytest.raises(Exception):
        oracle.sparse_motions(kpd, kps, 16, 16)


@pytest.mark.parametrize("name,cfg_name,audio", [("kp_tiny_b2", "tiny", False), ("kp_a_tiny_b3", "tiny", True),
                                                 ("kp_full_b2", "full", False), ("kp_a_full_b2", "full", True)])
def test_kp_detector_oracle_reproduces_reference_golden(name, cfg_name, audio):
    """KPDetector / KPDetector_a heads, pinned to the real reference by tools/make_golden.py."""
    from eamm_b200.config import get_kp_config
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    blob = np.load(os.path.join(GOLD, name + ".npz"))
    batch, size, _ = [int(v) for v in blob["meta"]]
    cfg = get_kp_config(cfg_name, audio=audio)
    sd = synth.make_kp_state_dict(cfg, seed=3 if audio else 2)
    x = synth.make_kp_inputs(cfg, batch, size, audio)
    np.testing.assert_allclose(np.array([x.double().sum()]), blob["in_checksum"], rtol=0, atol=0)
    got = (oracle.kp_detector_a_forward if audio else oracle.kp_detector_forward)(sd, cfg, x)
    for k in ("value", "heatmap", "jacobian"):
        a = got[k].numpy()
        sub = a[..., ::2, ::2] if (k == "heatmap" and cfg_name == "full") else a
        np.testing.assert_allclose(sub, blob[k], rtol=0, atol=2e-6, err_msg=k)
    assert np.allclose(got["heatmap"].sum((2, 3)).numpy(), 1.0, atol=1e-5)       # each heatmap is a distribution


@pytest.mark.parametrize("name", ["kp_glue_emo_t12", "kp_glue_plain_t40"])
def test_kp_glue_oracle_reproduces_reference_golden(name):
    """One-Euro smoothing + emotion rows + normalize_kp over a clip (demo.py:228-278)."""
    from oracle import kp_glue
    blob = np.load(os.path.join(GOLD, name + ".npz"))
    T, with_emo = [int(v) for v in blob["meta"]]
    drv, emo, src, init = synth.make_clip_inputs(T=T)
    v, j = kp_glue.clip_glue(drv["value"], drv["jacobian"], emo["value"] if with_emo else None,
                             emo["jacobian"] if with_emo else None, src, init, movement_scale=float(blob["scale"][0]))
    np.testing.assert_array_equal(v.numpy(), blob["value"])
    np.testing.assert_array_equal(j.numpy(), blob["jacobian"])
    # the filter starts from the first frame and the smoothed track stays within the raw track's range
    assert v.shape == (T, 10, 2) and j.shape == (T, 10, 2, 2)


@pytest.mark.parametrize("name", ["at_b2_t3", "at_b1_t6"])
def test_at_net2_oracle_reproduces_reference_golden(name):
    """AT_net2 (MFCC conv encoder + pose MLP + image DownBlocks -> 3-layer LSTM -> ConvTranspose
    stack), pinned to the real reference by tools/make_golden.py."""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    blob = np.load(os.path.join(GOLD, name + ".npz"))
    B, T = [int(v) for v in blob["meta"]]
    sd = synth.make_at_state_dict()
    img, mfcc, pose = synth.make_at_inputs(B, T)
    chk = np.array([img.double().sum(), mfcc.double().sum(), pose.double().sum()])
    np.testing.assert_allclose(chk, blob["in_checksum"], rtol=0, atol=0)
    taps = {}
    got = oracle.at_net2_forward(sd, img, mfcc, pose, 1.6, taps)
    assert got.shape == (B, T, 35, 64, 64)
    np.testing.assert_allclose(got.numpy()[..., ::4, ::4], blob["out"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(taps["lstm_out"].numpy(), blob["lstm_out"], rtol=0, atol=1e-6)
    # the step-by-step restatement of the LSTM recurrence agrees with torch's fused op
    np.testing.assert_allclose(oracle.lstm_explicit(taps["lstm_input"], sd).numpy(), blob["lstm_out"], rtol=0, atol=1e-6)
    # frames of one clip differ (the audio drives them) and the first frame depends on nothing later
    assert float((got[:, 0] - got[:, 1]).abs().mean()) > 1e-2
    head = oracle.at_net2_forward(sd, img, mfcc[:, :1], pose[:, :1], 1.6)
    np.testing.assert_allclose(head[:, 0].numpy(), got[:, 0].numpy(), rtol=0, atol=5e-6)


def lrw_clip_case():
    """BASELINE.json configs[4] fixture: real MFCC / pose / source frame of the reference's LRW sample, tiled to 300 windows."""
    from eamm_b200.clip import clip_inputs_from_windows
    blob = np.load(os.path.join(GOLD, "clip_lrw_t300.npz"))
    T = int(blob["meta"][0])
    img = (torch.from_numpy(blob["pixels_u8"]).float() / 255.0).permute(2, 0, 1).unsqueeze(0).contiguous()
    mfcc, pose = clip_inputs_from_windows(blob["mfcc13"], blob["pose7"], T)
    return blob, T, img, mfcc, pose


def test_oracle_chain_reproduces_the_reference_demo_path_on_real_mfcc():
    """configs[4]: AT_net2 -> KPDetector_a -> One-Euro + normalize_kp -> generator on the LRW sample, against the
    fixture the reference's own modules produced (tools/make_golden.py run_clip_case).  All 300 frames of keypoints;
    three generator frames (the CPU generator costs ~0.2 s per frame)."""
    from eamm_b200 import clip
    from eamm_b200.config import get_kp_config
    from oracle import kp_glue
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    blob, T, img, mfcc, pose = lrw_clip_case()
    assert mfcc.shape == (1, 300, 28, 12) and pose.shape == (1, 300, 6)
    cfg, kcfg, acfg = get_config("full"), get_kp_config("full"), get_kp_config("full", audio=True)
    deco = oracle.at_net2_forward(synth.make_at_state_dict(), img, mfcc, pose, 1.6)
    assert abs(float(deco.abs().mean()) - float(blob["deco_absmean"][0])) <= 1e-6
    src = oracle.kp_detector_forward(synth.make_kp_state_dict(kcfg, seed=2), kcfg, img)
    drv = oracle.kp_detector_a_forward(synth.make_kp_state_dict(acfg, seed=3), acfg, deco[0])
    init = {k: drv[k][:1] for k in ("value", "jacobian")}
    scale = clip.movement_scale(src, init)
    assert abs(scale - float(blob["scale"][0])) <= 1e-6
    nv, nj = kp_glue.clip_glue(drv["value"], drv["jacobian"], None, None, src, init, movement_scale=scale)
    np.testing.assert_allclose(nv.numpy(), blob["kp_value"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(nj.numpy(), blob["kp_jacobian"], rtol=0, atol=1e-5)
    idx = [0, 137, 299]
    out = oracle.generator_forward(synth.make_state_dict(cfg, seed=0), cfg, img.expand(3, -1, -1, -1).contiguous(),
                                   {"value": nv[idx], "jacobian": nj[idx]},
                                   {k: src[k].expand(3, *src[k].shape[1:]).contiguous() for k in ("value", "jacobian")})
    frames = oracle.frames_u8(out["prediction"])[:, ::8, ::8].numpy()
    assert np.abs(frames.astype(int) - blob["frames_u8_s8"][idx].astype(int)).max() <= 1      # batch-of-3 vs batch-of-1 rounding
    np.testing.assert_allclose(out["prediction"].double().sum((1, 2, 3)).numpy(), blob["frame_sums"][idx], rtol=1e-6)
