#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (authoring container only).

Imports /root/reference (read-only), loads the seeded synthetic state dict of eamm_b200.synth into
the real ``OcclusionAwareGenerator`` (strict=True, which also pins the 196-key layout), runs it on
CPU fp32 and
  1. asserts the restatement in oracle/eamm_oracle.py reproduces every output bit-for-bit,
  2. writes the outputs as fixtures: full tensors for the tiny config, strided sub-samples plus
     float64 checksums for the full 256x256 config (kept small on purpose).

Usage:  python tools/make_golden.py            (needs /root/reference; never runs on the GPU box)
        python tools/make_golden.py --verify   regenerate every fixture into a scratch directory (same oracle == reference
                                               assertions) and check that the committed tests/golden/*.npz hold the same arrays
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")
warnings.filterwarnings("ignore")

from eamm_b200.config import get_config          # noqa: E402
from eamm_b200 import synth                      # noqa: E402
from oracle import eamm_oracle as oracle         # noqa: E402

from eamm_b200.config import get_kp_config     # noqa: E402
from modules.generator import OcclusionAwareGenerator  # noqa: E402  (the reference)
from modules.keypoint_detector import KPDetector, KPDetector_a  # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "tests", "golden")        # --verify regenerates into a scratch directory instead
KEYS = ["mask", "sparse_deformed", "occlusion_map", "deformed", "prediction"]
STRIDES = {"mask": 4, "sparse_deformed": 4, "occlusion_map": 4, "deformed": 8, "prediction": 8, "deformation": 4}


def run_case(name, cfg_name, batch, size, with_jacobian=True, shared_source=False, full=True):
    cfg = get_config(cfg_name)
    sd = synth.make_state_dict(cfg, seed=0)
    ref = OcclusionAwareGenerator(**cfg).eval()
    missing = ref.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    src, kpd, kps = synth.make_inputs(batch, cfg, size=size, seed=1, with_jacobian=with_jacobian,
                                      shared_source=shared_source)
    has_dm = ref.dense_motion_network is not None
    with torch.no_grad():
        want = ref(src, kp_driving=kpd, kp_source=kps)
        want_dm = ref.dense_motion_network(source_image=src, kp_driving=kpd, kp_source=kps) if has_dm else {}
    got = oracle.generator_forward(sd, cfg, src, kpd, kps)
    got_dm = oracle.dense_motion_forward(sd, cfg, src, kpd, kps) if has_dm else {}
    assert set(want) == set(got), (sorted(want), sorted(got))
    for k in want:
        assert torch.equal(want[k], got[k]), f"{name}: oracle != reference on {k}"
    for k in want_dm:
        assert torch.equal(want_dm[k], got_dm[k]), f"{name}: oracle != reference on dense_motion.{k}"
    want = dict(want)
    if has_dm:
        want["deformation"] = want_dm["deformation"]
    blob = {"meta": np.array([batch, size, int(with_jacobian), int(shared_source)], dtype=np.int64)}
    blob["in_checksum"] = np.array([src.double().sum(), kpd["value"].double().sum(), kps["value"].double().sum()])
    for k, v in want.items():
        a = v.numpy()
        blob["sum_" + k] = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
        if full:
            blob[k] = a
        else:
            s = STRIDES[k]
            blob[k] = a[..., ::s, ::s].copy() if k != "deformation" else a[:, ::s, ::s, :].copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    stats = {k: (float(v.min()), float(v.max()), float(v.mean())) for k, v in want.items()}
    print(name, "ok ->", path, os.path.getsize(path) // 1024, "KiB")
    for k, s in stats.items():
        print("   ", k, "min/max/mean = %.4f %.4f %.4f" % s)


NATURAL = ["anne.png", "mona.png", "14.png", "jake4.png"]
NATURAL_STRIDES = {"mask": 2, "sparse_deformed": 2, "occlusion_map": 2, "deformed": 4, "prediction": 4, "deformation": 2}


def run_natural_case(name):
    """Natural 256x256 source images (the reference's own demo assets, /root/reference/test/image/*.png, read the way
    demo.py:476-478 feeds them: RGB / 255 as float32) with the synthetic keypoint recipe.  On a natural image the flow
    error -> intensity amplification of `deformed` is what a real run sees (warped outputs <= 1e-3).
    The fixture carries the uint8 pixels, so the GPU box needs nothing from the reference tree."""
    import cv2
    cfg = get_config("full")
    sd = synth.make_state_dict(cfg, seed=0)
    ref = OcclusionAwareGenerator(**cfg).eval()
    ref.load_state_dict(sd, strict=True)
    px = np.stack([cv2.cvtColor(cv2.imread("/root/reference/test/image/" + f, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
                   for f in NATURAL])                                  # [B,256,256,3] uint8
    src = natural_source(px)
    _, kpd, kps = synth.make_inputs(len(NATURAL), cfg, size=256, seed=1)
    with torch.no_grad():
        want = dict(ref(src, kp_driving=kpd, kp_source=kps))
        want["deformation"] = ref.dense_motion_network(source_image=src, kp_driving=kpd, kp_source=kps)["deformation"]
    taps = {}
    got = oracle.generator_forward(sd, cfg, src, kpd, kps, taps=taps)
    for k in KEYS:
        assert torch.equal(want[k], got[k]), f"{name}: oracle != reference on {k}"
    assert torch.equal(want["deformation"], taps["deformation"])
    blob = {"meta": np.array([len(NATURAL), 256, 1, 0], dtype=np.int64), "pixels_u8": px}
    for k, v in want.items():
        a = v.numpy()
        blob["sum_" + k] = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
        s = NATURAL_STRIDES[k]                                         # denser than the synthetic cases: 4 frames only
        blob[k] = a[..., ::s, ::s].copy() if k != "deformation" else a[:, ::s, ::s, :].copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print(name, "ok ->", path, os.path.getsize(path) // 1024, "KiB")


def natural_source(px):
    """uint8 [B,H,W,3] RGB -> float32 [B,3,H,W] in [0,1] (skimage img_as_float32 = x / 255, demo.py:476-478, :198)."""
    return (torch.from_numpy(px).float() / 255.0).permute(0, 3, 1, 2).contiguous()


def run_kp_case(name, cfg_name, batch, size, audio):
    cfg = get_kp_config(cfg_name, audio=audio)
    sd = synth.make_kp_state_dict(cfg, seed=3 if audio else 2)
    ref = (KPDetector_a if audio else KPDetector)(**cfg).eval()
    res = ref.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    x = synth.make_kp_inputs(cfg, batch, size, audio)
    with torch.no_grad():
        want = ref(x)
    got = (oracle.kp_detector_a_forward if audio else oracle.kp_detector_forward)(sd, cfg, x)
    blob = {"meta": np.array([batch, size, int(audio)], dtype=np.int64), "in_checksum": np.array([x.double().sum()])}
    for k in ("value", "heatmap", "jacobian"):
        assert torch.equal(want[k], got[k]), f"{name}: oracle != reference on {k}"
        a = want[k].numpy()
        blob["sum_" + k] = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
        blob[k] = a[..., ::2, ::2].copy() if (k == "heatmap" and cfg_name == "full") else a
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print(name, "ok ->", path, os.path.getsize(path) // 1024, "KiB", "value absmax %.3f" % want["value"].abs().max())


def run_glue_case(name, T, with_emo):
    """execute the reference's own OneEuroFilter (filter1.py) and normalize_kp (demo.py)
    source text on a seeded clip, following demo.py:228-278, and pin oracle/kp_glue.py to it bit-for-bit."""
    from scipy.spatial import ConvexHull
    from oracle import kp_glue
    txt = open("/root/reference/filter1.py").read()
    ns = {"np": np}
    exec(txt[txt.index("class LowPassFilter"):], ns)
    one_euro = ns["OneEuroFilter"]
    demo = open("/root/reference/demo.py").read()
    a = demo.index("def normalize_kp(")
    ns2 = {"np": np, "torch": torch, "ConvexHull": ConvexHull}
    exec(demo[a:demo.index("\ndef ", a + 10)], ns2)
    drv, emo, src, init = synth.make_clip_inputs(T=T)
    kp_all = [{k: v[t:t + 1].clone() for k, v in drv.items()} for t in range(T)]
    emo_all = [{k: v[t:t + 1].clone() for k, v in emo.items()} for t in range(T)]
    if with_emo:                                                     # demo.py:231-238
        fv, fj = one_euro(mincutoff=1, beta=0.2, dcutoff=1.0, freq=100), one_euro(mincutoff=1, beta=0.2, dcutoff=1.0, freq=100)
        for j in range(T):
            emo_all[j]["value"] = fv.process(emo_all[j]["value"] * 100) / 100
            emo_all[j]["jacobian"] = fj.process(emo_all[j]["jacobian"] * 100) / 100
    fv, fj = one_euro(mincutoff=0.05, beta=8, dcutoff=1.0, freq=100), one_euro(mincutoff=0.05, beta=8, dcutoff=1.0, freq=100)
    for j in range(T):                                               # demo.py:241-248
        kp_all[j]["value"] = fv.process(kp_all[j]["value"] * 10) / 10
        kp_all[j]["jacobian"] = fj.process(kp_all[j]["jacobian"] * 10) / 10
    rv, rj = [], []
    for t in range(T):                                               # demo.py:251-278
        kd, em = kp_all[t], emo_all[t]
        if with_emo:
            kd["value"][:, 1] = kd["value"][:, 1] + em["value"][:, 0] * 0.2
            kd["jacobian"][:, 1] = kd["jacobian"][:, 1] + em["jacobian"][:, 0] * 0.2
            kd["value"][:, 4] = kd["value"][:, 4] + em["value"][:, 1]
            kd["jacobian"][:, 4] = kd["jacobian"][:, 4] + em["jacobian"][:, 1]
            kd["value"][:, 6] = kd["value"][:, 6] + em["value"][:, 2]
            kd["jacobian"][:, 6] = kd["jacobian"][:, 6] + em["jacobian"][:, 2]
        n = ns2["normalize_kp"](kp_source=src, kp_driving=kd, kp_driving_initial=init, use_relative_movement=True,
                                use_relative_jacobian=True, adapt_movement_scale=True)
        rv.append(n["value"]); rj.append(n["jacobian"])
    rv, rj = torch.cat(rv), torch.cat(rj)
    scale = float(np.sqrt(ConvexHull(src["value"][0].numpy()).volume) / np.sqrt(ConvexHull(init["value"][0].numpy()).volume))
    ov, oj = kp_glue.clip_glue(drv["value"], drv["jacobian"], emo["value"] if with_emo else None,
                               emo["jacobian"] if with_emo else None, src, init, movement_scale=scale, relative=True)
    assert torch.equal(ov, rv) and torch.equal(oj, rj), name + ": oracle != reference"
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array([T, int(with_emo)]), scale=np.array([scale]), value=rv.numpy(), jacobian=rj.numpy())
    print(name, "ok ->", path, "scale %.6f" % scale)


def run_at_case(name, B, T):
    """the real AT_net2 (util.py:514-613) on CPU.  Its forward hard-codes `.cuda()` for the
    initial LSTM state (util.py:581-582); this script (not the reference) makes that call an identity."""
    from modules.util import AT_net2
    torch.Tensor.cuda = lambda self, *a, **k: self
    sd = synth.make_at_state_dict()
    ref = AT_net2().eval()
    res = ref.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.startswith("generator.") for k in res.missing_keys)
    img, mfcc, pose = synth.make_at_inputs(B, T)
    with torch.no_grad():
        want = ref(img, mfcc, pose, "cnn", 1.6)
    taps = {}
    got = oracle.at_net2_forward(sd, img, mfcc, pose, 1.6, taps)
    assert torch.equal(want, got), name + ": oracle != reference"
    a = want.numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array([B, T]), out=a[..., ::4, ::4].copy(), lstm_out=taps["lstm_out"].numpy(),
                        in_checksum=np.array([img.double().sum(), mfcc.double().sum(), pose.double().sum()]),
                        sum_out=np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()]))
    print(name, "ok ->", path, os.path.getsize(path) // 1024, "KiB", "abs mean %.4f" % np.abs(a).mean())


def run_clip_case(name, T=300):
    """BASELINE.json configs[4]: the demo.py path on REAL MFCC -- the in-tree LRW sample
    (/root/reference/dataset/LRW/{MFCC,Pose,Image}/ABOUT/ABOUT_00001*) -> AT_net2 -> KPDetector_a per frame -> One-Euro
    + normalize_kp -> generator per frame (demo.py:345, :206-281), every stage being the reference's own module or
    source text, with the seeded synthetic weights (no checkpoint ships).  The fixture carries the raw MFCC / pose /
    source pixels, the normalised keypoints and the frames as uint8 (stride-8 sub-sample) + per-frame checksums."""
    import cv2
    from scipy.spatial import ConvexHull
    from modules.util import AT_net2
    torch.Tensor.cuda = lambda self, *a, **k: self                       # AT_net2.forward hard-codes .cuda() (util.py:581)
    cfg, kcfg, acfg = get_config("full"), get_kp_config("full"), get_kp_config("full", audio=True)
    sd, ksd, asd, atsd = (synth.make_state_dict(cfg, seed=0), synth.make_kp_state_dict(kcfg, seed=2),
                          synth.make_kp_state_dict(acfg, seed=3), synth.make_at_state_dict())
    gen = OcclusionAwareGenerator(**cfg).eval(); gen.load_state_dict(sd, strict=True)
    det = KPDetector(**kcfg).eval(); det.load_state_dict(ksd, strict=True)
    det_a = KPDetector_a(**acfg).eval(); det_a.load_state_dict(asd, strict=True)
    at = AT_net2().eval(); at.load_state_dict(atsd, strict=False)
    base = "/root/reference/dataset/LRW/"
    mfcc13 = np.load(base + "MFCC/ABOUT/ABOUT_00001.npy")              # (30, 28, 13) float64
    pose7 = np.load(base + "Pose/ABOUT/ABOUT_00001.npy")               # (29, 7) float64
    px = cv2.cvtColor(cv2.imread(base + "Image/ABOUT/ABOUT_00001/0.png", cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    img = natural_source(px[None])
    from eamm_b200.clip import clip_inputs_from_windows
    mfcc, pose = clip_inputs_from_windows(mfcc13, pose7, T)
    txt = open("/root/reference/filter1.py").read()
    ns = {"np": np}
    exec(txt[txt.index("class LowPassFilter"):], ns)
    demo = open("/root/reference/demo.py").read()
    a = demo.index("def normalize_kp(")
    ns2 = {"np": np, "torch": torch, "ConvexHull": ConvexHull}
    exec(demo[a:demo.index("\ndef ", a + 10)], ns2)
    with torch.no_grad():
        deco = at(img, mfcc, pose, "cnn", 1.6)                           # demo.py:345
        kp_source = det(img)                                             # demo.py:206
        kp_init = det_a(deco[:, 0])
        kp_all = [det_a(deco[:, t]) for t in range(T)]                   # demo.py:219
        fv = ns["OneEuroFilter"](mincutoff=0.05, beta=8, dcutoff=1.0, freq=100)
        fj = ns["OneEuroFilter"](mincutoff=0.05, beta=8, dcutoff=1.0, freq=100)
        for j in range(T):                                               # demo.py:241-248
            kp_all[j]["value"] = fv.process(kp_all[j]["value"] * 10) / 10
            kp_all[j]["jacobian"] = fj.process(kp_all[j]["jacobian"] * 10) / 10
        nv, nj, frames, sums = [], [], [], []
        for t in range(T):                                               # demo.py:251-281
            kn = ns2["normalize_kp"](kp_source=kp_source, kp_driving=kp_all[t], kp_driving_initial=kp_init,
                                     use_relative_movement=True, use_relative_jacobian=True, adapt_movement_scale=True)
            out = gen(img, kp_source=kp_source, kp_driving=kn)
            nv.append(kn["value"]); nj.append(kn["jacobian"])
            p = out["prediction"]
            frames.append(oracle.frames_u8(p)[0, ::8, ::8].numpy())
            sums.append(float(p.double().sum()))
            if t % 50 == 0:
                print("  frame", t, flush=True)
    nv, nj = torch.cat(nv).numpy(), torch.cat(nj).numpy()
    # the oracle chain (what the GPU box can re-run) reproduces the reference chain
    from oracle import kp_glue
    o_deco = oracle.at_net2_forward(atsd, img, mfcc, pose, 1.6)
    assert torch.equal(o_deco, deco)
    o_src = oracle.kp_detector_forward(ksd, kcfg, img)
    o_drv = oracle.kp_detector_a_forward(asd, acfg, o_deco[0])
    o_init = {k: o_drv[k][:1] for k in ("value", "jacobian")}
    scale = float(np.sqrt(ConvexHull(o_src["value"][0].numpy()).volume) / np.sqrt(ConvexHull(o_init["value"][0].numpy()).volume))
    ov, oj = kp_glue.clip_glue(o_drv["value"], o_drv["jacobian"], None, None, o_src, o_init, movement_scale=scale)
    assert np.abs(ov.numpy() - nv).max() <= 1e-6 and np.abs(oj.numpy() - nj).max() <= 1e-5, "oracle chain != reference chain"
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array([T]), mfcc13=mfcc13.astype(np.float32), pose7=pose7.astype(np.float32),
                        pixels_u8=px, kp_value=nv, kp_jacobian=nj, frames_u8_s8=np.stack(frames),
                        frame_sums=np.array(sums), scale=np.array([scale]),
                        deco_absmean=np.array([float(deco.abs().mean())]))
    print(name, "ok ->", path, os.path.getsize(path) // 1024, "KiB; kp motion over the clip: value std %.4f" %
          float(nv.std(axis=0).mean()))


def verify_committed(scratch):
    """Compare the freshly generated fixtures in `scratch` with the committed ones, array by array (bit-exact)."""
    gold = os.path.join(ROOT, "tests", "golden")
    names = sorted(f for f in os.listdir(gold) if f.endswith(".npz"))
    bad = [f for f in names if not os.path.exists(os.path.join(scratch, f))]
    for f in names:
        if f in bad:
            continue
        a, b = np.load(os.path.join(gold, f), allow_pickle=True), np.load(os.path.join(scratch, f), allow_pickle=True)
        same = set(a.files) == set(b.files) and all(
            a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes() for k in a.files)
        if not same:
            bad.append(f)
    print("verify: %d committed fixtures, %d regenerate bit-identically from the reference%s"
          % (len(names), len(names) - len(bad), "" if not bad else "; DIFFERENT: " + ", ".join(bad)))
    return 1 if bad else 0


if __name__ == "__main__":
    torch.set_num_threads(8)
    verify = "--verify" in sys.argv[1:]
    if verify:
        import tempfile
        sys.argv.remove("--verify")
        OUT = tempfile.mkdtemp(prefix="eamm_golden_")
    only = sys.argv[1:]
    if only:                                   # python tools/make_golden.py natural_b4 ...  (regenerate selected fixtures)
        table = {"natural_b4": lambda: run_natural_case("natural_b4"), "clip_lrw_t300": lambda: run_clip_case("clip_lrw_t300"),
                 "tiny_sf05_b2": lambda: run_case("tiny_sf05_b2", "tiny_sf05", 2, 64),
                 "tiny_sf1_b2": lambda: run_case("tiny_sf1_b2", "tiny_sf1", 2, 64),
                 "tiny_nodm_b2": lambda: run_case("tiny_nodm_b2", "tiny_nodm", 2, 64)}
        for nm in only:
            table[nm]()
        sys.exit(0)
    run_natural_case("natural_b4")
    run_clip_case("clip_lrw_t300")
    run_case("tiny_b2", "tiny", 2, 64)
    run_case("tiny_b3_nojac", "tiny", 3, 64, with_jacobian=False)
    # constructor corners no shipped config uses: flow / occlusion resize, no anti-alias module, no dense-motion network
    run_case("tiny_sf05_b2", "tiny_sf05", 2, 64)
    run_case("tiny_sf1_b2", "tiny_sf1", 2, 64)
    run_case("tiny_nodm_b2", "tiny_nodm", 2, 64)
    run_case("full_b2", "full", 2, 256, full=False)
    run_case("full_b3_shared", "full", 3, 256, shared_source=True, full=False)
    # BASELINE.json configs[0]: one 256x256 source + 16 synthetic kp/jacobian frames
    run_case("full_b16_shared", "full", 16, 256, shared_source=True, full=False)
    run_kp_case("kp_tiny_b2", "tiny", 2, 64, audio=False)
    run_kp_case("kp_a_tiny_b3", "tiny", 3, 64, audio=True)
    run_kp_case("kp_full_b2", "full", 2, 256, audio=False)
    run_kp_case("kp_a_full_b2", "full", 2, 256, audio=True)
    run_glue_case("kp_glue_emo_t12", 12, True)
    run_glue_case("kp_glue_plain_t40", 40, False)
    run_at_case("at_b2_t3", 2, 3)
    run_at_case("at_b1_t6", 1, 6)
    if verify:
        sys.exit(verify_committed(OUT))
