"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the drop-in
modules and directly through the C ABI, against the CPU oracle and the committed golden fixtures.

Tolerances (max-abs):
  fp32_simt : 2e-5 everywhere (true fp32 CUDA-core convs), warped images 2e-4
  fp32      : prediction/mask/occlusion/deformation 1e-4, sparse_deformed 1e-4, deformed 3e-3 on the synthetic
              white-noise source (it turns a 1e-5 flow error into 1e-3 of intensity, DESIGN.md "Numerics") and the
              survey's 1e-3 on natural images (test_natural_images_match_reference_golden); the same bounds for
              fp32_bf16x3 (every layer on the 3-pass bf16 hi/lo scheme instead of fp16 + fp8 in the bottleneck)
  fp16      : prediction/mask/occlusion 1e-2 and PSNR >= 50 dB (SURVEY 8(c)'s reduced-precision bound: one-pass
              convs with fp16 operands; flow-sensitive outputs are not pinned)
  bf16      : prediction 3e-2, mask/occlusion 3e-2 (the literal "bf16 conv" wording of BASELINE configs[2]: 8
              significant bits, measured 1.1e-2 on the white-noise source -- kept as a mode, not the recommended one)
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from eamm_b200 import get_config, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {
    "fp32_simt": {"prediction": 2e-5, "mask": 2e-5, "occlusion_map": 2e-5, "deformation": 2e-5,
                  "sparse_deformed": 2e-5, "deformed": 2e-4},
    "fp32": {"prediction": 1e-4, "mask": 1e-4, "occlusion_map": 1e-4, "deformation": 1e-4,
             "sparse_deformed": 1e-4, "deformed": 3e-3},
    "fp16": {"prediction": 1e-2, "mask": 1e-2, "occlusion_map": 1e-2, "sparse_deformed": 1e-4},
    "bf16": {"prediction": 3e-2, "mask": 3e-2, "occlusion_map": 3e-2, "sparse_deformed": 1e-4},
}
TOL["fp32_bf16x3"] = TOL["fp32"]
STRIDES = {"mask": 4, "sparse_deformed": 4, "occlusion_map": 4, "deformed": 8, "prediction": 8, "deformation": 4}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


_GEN = {}


def generator(cfg_name, dev):
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    if cfg_name not in _GEN:
        cfg = get_config(cfg_name)
        g = OcclusionAwareGenerator(**cfg).eval()
        g.load_state_dict(synth.make_state_dict(cfg, seed=0), strict=True)
        _GEN[cfg_name] = (g.to(dev), cfg)
    return _GEN[cfg_name]


def to_dev(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


def run_ours(cfg_name, dev, precision, src, kpd, kps, shared=False):
    gen, cfg = generator(cfg_name, dev)
    gen.precision = precision
    s = src.to(dev)
    if shared:
        s = s[:1].expand(src.shape[0], -1, -1, -1)
    out = gen(s, kp_driving=to_dev(kpd, dev), kp_source=to_dev(kps, dev))
    out = dict(out)
    if gen._eng.dm is not None:
        out["deformation"] = gen._eng.last_dm["deformation"]
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items()}


def test_native_library_is_the_in_tree_build_and_device_is_sm100(dev):
    from eamm_b200 import _lib, build
    lib = _lib.load()
    assert os.path.dirname(build.LIBPATH).endswith(os.path.join("eamm_b200", "lib"))
    assert lib.eamm_device_ok(0) == 1


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32", "fp16", "bf16"])
@pytest.mark.parametrize("case,cfg_name", [("tiny_b2", "tiny"), ("tiny_b3_nojac", "tiny"),
                                           # constructor corners of the reference no shipped config uses: motion grid 2x the
                                           # feature grid (generator.py:53-56, :82-83), scale_factor 1 (dense_motion.py:82),
                                           # dense_motion_params=None (generator.py:20-24, :67)
                                           ("tiny_sf05_b2", "tiny_sf05"), ("tiny_sf1_b2", "tiny_sf1"),
                                           ("tiny_nodm_b2", "tiny_nodm")])
def test_tiny_config_matches_reference_golden(dev, precision, case, cfg_name):
    blob = np.load(os.path.join(GOLD, case + ".npz"))
    batch, size, jac, shared = [int(v) for v in blob["meta"]]
    cfg = get_config(cfg_name)
    src, kpd, kps = synth.make_inputs(batch, cfg, size=size, seed=1, with_jacobian=bool(jac))
    got = run_ours(cfg_name, dev, precision, src, kpd, kps)
    if cfg["dense_motion_params"] is None:
        assert set(got) == {"prediction"}
    for k, tol in TOL[precision].items():
        if k not in got:
            continue
        err = np.abs(got[k].numpy() - blob[k]).max()
        assert err <= tol, "%s %s %s: max-abs %.3e > %.1e" % (case, precision, k, err, tol)


@pytest.mark.parametrize("precision", ["fp32", "fp32_bf16x3", "fp16", "bf16"])
@pytest.mark.parametrize("case", ["full_b2", "full_b3_shared", "full_b16_shared"])   # the last = BASELINE configs[0]
def test_full_config_matches_reference_golden(dev, precision, case):
    blob = np.load(os.path.join(GOLD, case + ".npz"))
    batch, size, jac, shared = [int(v) for v in blob["meta"]]
    cfg = get_config("full")
    src, kpd, kps = synth.make_inputs(batch, cfg, size=size, seed=1, shared_source=bool(shared))
    got = run_ours("full", dev, precision, src, kpd, kps, shared=bool(shared))
    for k, tol in TOL[precision].items():
        a = got[k].numpy()
        s = STRIDES[k]
        sub = a[:, ::s, ::s, :] if k == "deformation" else a[..., ::s, ::s]
        err = np.abs(sub - blob[k]).max()
        print("%s %s %-16s max-abs %.3e (tol %.1e)" % (case, precision, k, err, tol))
        assert err <= tol, "%s %s %s: max-abs %.3e > %.1e" % (case, precision, k, err, tol)
        if precision == "fp16" and k == "prediction":
            assert psnr(torch.from_numpy(sub), torch.from_numpy(blob[k])) >= 50.0
        if precision.startswith("fp32"):   # checksum of the whole tensor, not just the stored sub-sample
            tot = a.astype(np.float64).sum()
            assert abs(tot - blob["sum_" + k][0]) <= tol * a.size * 0.3 + 1e-3, k


def psnr(a, b):
    return float(-10.0 * torch.log10(((a.double() - b.double()) ** 2).mean()))


# Natural source images (the reference's own demo assets; pixels carried by the fixture).  Here the survey's tolerance
# for the warped outputs holds as stated (deformed / sparse_deformed <= 1e-3): a natural image has none of
# the white-noise gradient that turns a 1e-5 flow error into 1e-3 of intensity in the synthetic cases.
NATURAL_TOL = {
    "fp32": {"prediction": 1e-4, "mask": 1e-4, "occlusion_map": 1e-4, "deformation": 1e-4, "sparse_deformed": 1e-4,
             "deformed": 1e-3},
    "fp16": {"prediction": 1e-2, "mask": 1e-2, "occlusion_map": 1e-2, "sparse_deformed": 1e-4},
}


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_natural_images_match_reference_golden(dev, precision):
    from test_oracle_golden import natural_case, natural_subsample
    blob, cfg, src, kpd, kps = natural_case()
    got = run_ours("full", dev, precision, src, kpd, kps)
    for k, tol in NATURAL_TOL[precision].items():
        err = np.abs(natural_subsample(k, got[k].numpy()) - blob[k]).max()
        print("natural images %s %-16s max-abs %.3e (tol %.1e)" % (precision, k, err, tol))
        assert err <= tol, (precision, k, err)
    p = psnr(torch.from_numpy(natural_subsample("prediction", got["prediction"].numpy())), torch.from_numpy(blob["prediction"]))
    print("natural images %s prediction PSNR %.1f dB" % (precision, p))
    assert p >= (50.0 if precision == "fp16" else 80.0)


def test_full_size_batch32_properties_and_batch_invariance(dev):
    """BASELINE configs[1] size: structural properties + frame i of the batch == the same frame run alone."""
    cfg = get_config("full")
    src, kpd, kps = synth.make_inputs(32, cfg, size=256, seed=7)
    got = run_ours("full", dev, "fp32", src, kpd, kps)
    assert all(torch.isfinite(v).all() for v in got.values())
    assert torch.allclose(got["mask"].sum(1), torch.ones(32, 64, 64), atol=1e-5)
    assert got["occlusion_map"].min() > 0 and got["occlusion_map"].max() < 1
    assert got["prediction"].min() > 0 and got["prediction"].max() < 1
    i = 17
    one = run_ours("full", dev, "fp32", src[i:i + 1], {k: v[i:i + 1] for k, v in kpd.items()},
                   {k: v[i:i + 1] for k, v in kps.items()})
    # Not bit-equal by design: the 8x8 ... 2x2 hourglass layers run split-K with a split factor chosen from the number
    # of frames in the launch (conv_tc.cu), so the fp32 summation grouping of those layers differs between B=32 and B=1.
    # Measured on B200: prediction 1.5e-5, mask 3.5e-6, deformed 5.8e-4 (the white-noise test image turns a 4e-6
    # flow difference into 6e-4 of intensity); bounds = half of the tolerances against the oracle.
    for k, tol in (("prediction", 5e-5), ("mask", 2e-5), ("deformed", 1.5e-3)):
        err = (one[k][0] - got[k][i]).abs().max().item()
        print("batch invariance %s max-abs %.3e" % (k, err))
        assert err <= tol, (k, err)
    # the oracle on ALL 32 frames and all outputs (the headline configuration; ~3 s of CPU on the bench box)
    from oracle import eamm_oracle as oracle
    sd = synth.make_state_dict(cfg, seed=0)
    taps = {}
    want = oracle.generator_forward(sd, cfg, src, kpd, kps, taps=taps)
    want["deformation"] = taps["deformation"]
    for k, tol in TOL["fp32"].items():
        err = (got[k] - want[k]).abs().max().item()
        print("configs[1] B=32 fp32 mode vs oracle: %-16s max-abs %.3e (tol %.1e)" % (k, err, tol))
        assert err <= tol, (k, err)


def test_shared_source_broadcast_equals_distinct_copies(dev):
    cfg = get_config("tiny")
    src, kpd, kps = synth.make_inputs(4, cfg, size=64, seed=11, shared_source=True)
    a = run_ours("tiny", dev, "fp32", src, kpd, kps, shared=True)
    b = run_ours("tiny", dev, "fp32", src, kpd, kps, shared=False)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_identity_keypoints_give_identity_flow_and_singular_jacobian_raises(dev):
    cfg = get_config("tiny")
    src, kpd, kps = synth.make_inputs(2, cfg, size=64, seed=13)
    got = run_ours("tiny", dev, "fp32_simt", src, kps, kps)          # driving == source
    from oracle import eamm_oracle as oracle
    ident = oracle.make_coordinate_grid(16, 16).view(1, 16, 16, 2).expand(2, -1, -1, -1)
    assert torch.allclose(got["deformation"], ident, atol=1e-5)
    bad = {k: v.clone() for k, v in kpd.items()}
    bad["jacobian"][1, 2] = 0.0
    with pytest.raises(torch.linalg.LinAlgError):
        run_ours("tiny", dev, "fp32_simt", src, bad, kps)


def test_dense_motion_module_alone_matches_oracle(dev):
    from eamm_b200.modules.generator import OcclusionAwareGenerator
    from oracle import eamm_oracle as oracle
    gen, cfg = generator("tiny", dev)
    gen.dense_motion_network.precision = "fp32"
    src, kpd, kps = synth.make_inputs(2, cfg, size=64, seed=17)
    out = gen.dense_motion_network(source_image=src.to(dev), kp_driving=to_dev(kpd, dev), kp_source=to_dev(kps, dev))
    want = oracle.dense_motion_forward(synth.make_state_dict(cfg, seed=0), cfg, src, kpd, kps)
    assert set(out) == {"sparse_deformed", "mask", "deformation", "occlusion_map"}
    for k in want:
        assert (out[k].cpu() - want[k]).abs().max() <= 1e-4, k


# ------------------------------------------------------------------ kernels through the C ABI
def test_cabi_warp_occlude_matches_grid_sample_times_occlusion(dev):
    from eamm_b200 import _lib as L
    from eamm_b200.engine import ActBuf, current_stream_ptr
    lib = L.load()
    g = torch.Generator().manual_seed(5)
    n, h, w, c = 3, 16, 16, 64
    feat = torch.randn(n, c, h, w, generator=g)
    grid = torch.rand(n, h, w, 2, generator=g) * 2.4 - 1.2
    grid[0, 0, 0] = torch.tensor([-1.0, -1.0]); grid[0, 0, 1] = torch.tensor([1.0, 1.0])   # corner unit vectors
    occ = torch.rand(n, 1, h, w, generator=g)
    want = F.grid_sample(feat, grid, align_corners=False) * occ
    fin = ActBuf(n, h, w, c, "f32", dev); fin.t.copy_(feat.permute(0, 2, 3, 1))
    fout = ActBuf(n, h, w, c, "f32", dev)
    gd, od = grid.to(dev), occ.to(dev)
    L.check(lib.eamm_warp_occlude(C.byref(fin.act()), gd.data_ptr(), od.data_ptr(), 0, 0, C.byref(fout.act()), None, None, None, None, current_stream_ptr()), "warp_occlude")
    torch.cuda.synchronize()
    assert (fout.to_float().cpu() - want).abs().max() <= 2e-6
    # NaN / inf flow coordinates give NaN like F.grid_sample (weights inf - inf); huge finite ones sample zeros
    grid3 = grid.clone()
    grid3[1, 2, 3, 0] = float("nan"); grid3[1, 2, 4, 1] = float("inf"); grid3[1, 2, 5, 0] = 1e30
    want3 = F.grid_sample(feat, grid3, align_corners=False) * occ
    L.check(lib.eamm_warp_occlude(C.byref(fin.act()), grid3.to(dev).data_ptr(), od.data_ptr(), 0, 0, C.byref(fout.act()), None, None,
                                  None, None, current_stream_ptr()), "warp_occlude")
    torch.cuda.synchronize()
    got3 = fout.to_float().cpu()
    assert torch.equal(torch.isnan(got3), torch.isnan(want3)) and torch.isnan(got3).sum() == 2 * c
    assert (torch.nan_to_num(got3) - torch.nan_to_num(want3)).abs().max() <= 2e-6
    # zero padding is bit-exact: a sample fully outside the image reads exactly 0
    grid2 = torch.full((n, h, w, 2), 3.0)
    L.check(lib.eamm_warp_occlude(C.byref(fin.act()), grid2.to(dev).data_ptr(), None, 0, 0, C.byref(fout.act()), None, None, None, None, current_stream_ptr()), "warp_occlude")
    torch.cuda.synchronize()
    assert fout.t.abs().max().item() == 0.0


def test_cabi_aa_downsample_and_warp_image_match_oracle(dev):
    from eamm_b200 import _lib as L
    from eamm_b200.engine import current_stream_ptr
    from oracle import eamm_oracle as oracle
    lib = L.load()
    g = torch.Generator().manual_seed(6)
    src = torch.rand(2, 3, 64, 64, generator=g)
    k2 = synth.aa_kernel(3)
    want = oracle.anti_alias_down(src, k2, 0.25)
    g1 = k2[0, 0].sum(1); g1 = (g1 / g1.sum()).to(dev)
    out = torch.zeros(2, 16, 16, 4, device=dev)
    sd = src.to(dev)
    L.check(lib.eamm_aa_downsample(sd.data_ptr(), 3 * 64 * 64, out.data_ptr(), 2, 64, 64, 4, g1.data_ptr(), 13,
                                   current_stream_ptr()), "aa")
    torch.cuda.synchronize()
    assert (out[..., :3].permute(0, 3, 1, 2).cpu() - want).abs().max() <= 1e-6
    flow = torch.rand(2, 16, 16, 2, generator=g) * 2.2 - 1.1
    want2 = oracle.deform_input(src, flow)
    got2 = torch.zeros(2, 3, 64, 64, device=dev)
    L.check(lib.eamm_warp_image(sd.data_ptr(), 3 * 64 * 64, flow.to(dev).data_ptr(), got2.data_ptr(), 2, 3, 64, 64, 16, 16,
                                current_stream_ptr()), "warp_image")
    torch.cuda.synchronize()
    assert (got2.cpu() - want2).abs().max() <= 1e-5


def test_cabi_conv_tc_matches_conv_simt_on_layer_shapes(dev):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "gpu_conv_check", os.path.join(os.path.dirname(GOLD), "..", "tools", "gpu_conv_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for idx in (1, 3, 5, 7, 8, 9, 10, 12, 13, 15, 17, 18, 19, 21, 22):
        assert mod.run_case(idx) == 0, mod.CASES[idx][0]
    # 112-column kx-in-N schemes and split-K: the case fails unless the planner really chose the scheme under test
    new = [i for i, c in enumerate(mod.CASES) if c[0].startswith(("kxw", "splitk", "f16", "mix", "ah "))]
    assert len(new) >= 32
    for idx in new:
        assert mod.run_case(idx) == 0, mod.CASES[idx][0]


@pytest.mark.parametrize("kind,cin,cout,N,H,W,flags", [
    ("3x3", 64, 128, 2, 32, 32, "rp"),      # DownBlock2d (util.py:915-920)
    ("3x3", 256, 256, 1, 16, 16, ""),       # ResBlock2d conv (util.py:876-878)
    ("up2", 256, 128, 2, 8, 8, "r"),        # UpBlock2d (util.py:895-900)
    ("up2", 44, 20, 1, 5, 7, "r"),          # odd map, ragged channels
    ("7x7", 4, 64, 1, 32, 32, "r"),         # `first` (generator.py:25)
    ("7x7", 108, 12, 2, 16, 16, ""),        # mask + occlusion logits (dense_motion.py:98,110)
    ("3x3", 1024, 1024, 3, 2, 2, "rp"),     # deepest hourglass encoder block
])
def test_conv_simt_anchor_matches_torch_conv2d(dev, kind, cin, cout, N, H, W, flags):
    """The fp32 CUDA-core convolution is the on-device anchor the tensor-core kernel is checked against
    (tools/gpu_conv_check.py); here the anchor itself meets F.conv2d (fp32, CPU) on every layer shape family."""
    from eamm_b200 import _lib as L
    from eamm_b200.engine import ActBuf, ConvLayer, current_stream_ptr
    lib = L.load()
    g = torch.Generator().manual_seed(cin * 131 + cout)
    ks = 7 if kind == "7x7" else 3
    w = (torch.rand(cout, cin, ks, ks, generator=g) * 2 - 1) * (3.0 / (cin * ks * ks)) ** 0.5
    b = (torch.rand(cout, generator=g) * 2 - 1) * 0.1
    x = torch.randn(N, cin, H, W, generator=g)
    xin = x
    if kind == "up2":
        xin = F.interpolate(x, scale_factor=2)
    want = F.conv2d(xin, w, b, padding=ks // 2)
    if "r" in flags:
        want = F.relu(want)
    if "p" in flags:
        want = F.avg_pool2d(want, 2)
    cs, co = (cin + 3) // 4 * 4, (cout + 3) // 4 * 4
    kk = {"3x3": L.CONV_3X3, "up2": L.CONV_UP2_3X3, "7x7": L.CONV_7X7}[kind]
    fl = (L.EPI_RELU if "r" in flags else 0) | (L.EPI_POOL2 if "p" in flags else 0)
    layer = ConvLayer("anchor", kk, fl, w.to(dev), b.to(dev), cs, 4, "simt")
    inb = ActBuf(N, H, W, cs, "f32", dev)
    inb.t[..., :cin].copy_(x.permute(0, 2, 3, 1))
    out = ActBuf(N, want.shape[2], want.shape[3], co, "f32", dev)
    layer.launch(lib, current_stream_ptr(), inb.act(), out=out.act())
    torch.cuda.synchronize()
    got = out.to_float(c=cout).cpu()
    err = (got - want).abs().max().item()
    assert err <= 2e-5 * max(1.0, want.abs().max().item()), (kind, cin, cout, err)


def test_cabi_rejects_bad_arguments_without_launching(dev):
    from eamm_b200 import _lib as L
    from eamm_b200.engine import ActBuf, current_stream_ptr
    lib = L.load()
    a = ActBuf(1, 6, 6, 64, "bf16", dev)          # 6x6 is not a power-of-two map
    args = L.ConvArgs()
    args.kind, args.flags, args.cin, args.cout = L.CONV_3X3, 0, 64, 16
    act = a.act()
    args.inp = C.pointer(act)
    w = torch.zeros(16, 9 * 64, dtype=torch.bfloat16, device=dev)
    b = torch.zeros(16, device=dev)
    args.weight, args.bias = w.data_ptr(), b.data_ptr()
    o = ActBuf(1, 6, 6, 16, "bf16", dev)
    oa = o.act()
    args.out = C.pointer(oa)
    assert lib.eamm_conv_tc(C.byref(args), current_stream_ptr()) == -5      # EAMM_ERR_UNSUPPORTED
    args.cin = 32
    assert lib.eamm_conv_tc(C.byref(args), current_stream_ptr()) < 0


def test_frame_pipeline_equals_direct_forward(dev):
    """Host-in / host-out pipelining (three streams) returns exactly what direct calls return."""
    from eamm_b200.pipeline import FramePipeline
    gen, cfg = generator("tiny", dev)
    gen.precision = "fp32"
    batches = [synth.make_inputs(3, cfg, size=64, seed=40 + i) for i in range(5)]
    want = [run_ours("tiny", dev, "fp32", *b)["prediction"] for b in batches]
    pipe = FramePipeline(gen, depth=2)
    outs = [torch.empty(3, 3, 64, 64).pin_memory() for _ in batches]
    pinned = [(s.pin_memory(), {k: v.pin_memory() for k, v in kd.items()}, {k: v.pin_memory() for k, v in ks.items()})
              for s, kd, ks in batches]
    for (s, kd, ks), o in zip(pinned, outs):
        pipe.submit(s, kd, ks, o)
    pipe.close()
    for o, w in zip(outs, want):
        assert torch.equal(o, w)


def test_frame_pipeline_reports_a_singular_jacobian_of_any_queued_batch(dev):
    """The device error flag is sticky across the non-strict forwards of a pipeline: a singular driving Jacobian in an
    EARLY batch still raises from drain() (the reference raises from torch.inverse, dense_motion.py:56)."""
    from eamm_b200.pipeline import FramePipeline
    gen, cfg = generator("tiny", dev)
    gen.precision = "fp32"
    batches = [synth.make_inputs(2, cfg, size=64, seed=90 + i) for i in range(4)]
    batches[1][1]["jacobian"][0, 1] = 0.0                  # batch 1 of 4 is the bad one
    pinned = [(s.pin_memory(), {k: v.pin_memory() for k, v in kd.items()}, {k: v.pin_memory() for k, v in ks.items()})
              for s, kd, ks in batches]
    outs = [torch.empty(2, 3, 64, 64).pin_memory() for _ in batches]
    pipe = FramePipeline(gen, depth=2)
    for (s, kd, ks), o in zip(pinned, outs):
        pipe.submit(s, kd, ks, o)
    with pytest.raises(torch.linalg.LinAlgError):
        pipe.drain()
    pipe.close()                                           # the flag was cleared by the failed drain
    assert gen.strict_errors
    out = run_ours("tiny", dev, "fp32", *batches[0])       # and the generator is usable afterwards
    assert torch.isfinite(out["prediction"]).all()


def test_graph_without_fixed_source_ignores_a_warm_source_cache(dev):
    """a caller that had `cache_source = True` must still get a graph that contains the encoder, so that a
    replay with a new source image renders the new identity."""
    from eamm_b200.graph import GraphedGenerator
    gen, cfg = generator("tiny", dev)
    gen.precision = "fp32"
    src_a, kpd, kps = synth.make_inputs(1, cfg, size=64, seed=95)
    src_b = synth.make_inputs(1, cfg, size=64, seed=96)[0]
    want_b = run_ours("tiny", dev, "fp32", src_b, kpd, kps)["prediction"]
    gen.cache_source = True
    try:
        sa = src_a.to(dev)
        gen(sa, kp_driving=to_dev(kpd, dev), kp_source=to_dev(kps, dev))          # warms the cache for src_a
        graphed = GraphedGenerator(gen, sa, to_dev(kpd, dev), to_dev(kps, dev), fixed_source=False)
        assert gen.cache_source is True                                          # restored for the caller
        got = graphed(src_b.to(dev), to_dev(kpd, dev), to_dev(kps, dev))["prediction"]
        torch.cuda.synchronize()
    finally:
        gen.cache_source = False
    assert torch.equal(got.cpu(), want_b)


def test_workspace_cache_is_bounded(dev):
    gen, cfg = generator("tiny", dev)
    gen.precision = "fp32"
    for b in (1, 2, 3, 4, 5):
        run_ours("tiny", dev, "fp32", *synth.make_inputs(b, cfg, size=64, seed=97))
    assert len(gen._eng.ws) <= 2 and len(gen._eng.dm.ws) <= 2


@pytest.mark.parametrize("size,batch", [(128, 3), (512, 1), (256, 5)])
def test_other_image_sizes_and_ragged_batches_match_oracle(dev, size, batch):
    """Full config at 128/512 px and a batch that does not fill the 128-pixel tiles of the small maps."""
    from oracle import eamm_oracle as oracle
    cfg = get_config("full")
    src, kpd, kps = synth.make_inputs(batch, cfg, size=size, seed=50 + size)
    got = run_ours("full", dev, "fp32", src, kpd, kps)
    want = oracle.generator_forward(synth.make_state_dict(cfg, seed=0), cfg, src, kpd, kps)
    for k, tol in (("prediction", 1e-4), ("mask", 1e-4), ("occlusion_map", 1e-4), ("sparse_deformed", 1e-4), ("deformed", 5e-3)):
        assert got[k].shape == want[k].shape, k
        err = (got[k] - want[k]).abs().max().item()
        assert err <= tol, "%dpx B=%d %s: %.3e" % (size, batch, k, err)


def test_empty_batch_returns_empty_tensors(dev):
    gen, cfg = generator("tiny", dev)
    gen.precision = "fp32"
    src, kpd, kps = synth.make_inputs(2, cfg, size=64, seed=3)
    out = gen(src[:0].to(dev), kp_driving={k: v[:0].to(dev) for k, v in kpd.items()},
              kp_source={k: v[:0].to(dev) for k, v in kps.items()})
    assert out["prediction"].shape == (0, 3, 64, 64) and out["mask"].shape == (0, 4, 16, 16)
    assert out["sparse_deformed"].shape == (0, 4, 3, 16, 16) and out["deformed"].shape == (0, 3, 64, 64)


# ------------------------------------------------------------------ keypoint heads
@pytest.mark.parametrize("precision,tol_v,tol_h", [("fp32_simt", 2e-5, 2e-5), ("fp32", 2e-4, 1e-4)])
@pytest.mark.parametrize("name,cfg_name,audio", [("kp_tiny_b2", "tiny", False), ("kp_a_tiny_b3", "tiny", True),
                                                 ("kp_full_b2", "full", False), ("kp_a_full_b2", "full", True)])
def test_kp_detector_heads_match_reference_golden(dev, precision, tol_v, tol_h, name, cfg_name, audio):
    from eamm_b200.config import get_kp_config
    from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
    blob = np.load(os.path.join(GOLD, name + ".npz"))
    batch, size, _ = [int(v) for v in blob["meta"]]
    cfg = get_kp_config(cfg_name, audio=audio)
    det = (KPDetector_a if audio else KPDetector)(**cfg).eval()
    det.load_state_dict(synth.make_kp_state_dict(cfg, seed=3 if audio else 2), strict=True)
    det = det.to(dev)
    det.precision = precision
    out = det(synth.make_kp_inputs(cfg, batch, size, audio).to(dev))
    torch.cuda.synchronize()
    assert set(out) == {"value", "heatmap", "jacobian"}
    hm = out["heatmap"].cpu().numpy()
    sub = hm[..., ::2, ::2] if cfg_name == "full" else hm
    assert np.abs(sub - blob["heatmap"]).max() <= tol_h
    assert np.abs(out["value"].cpu().numpy() - blob["value"]).max() <= tol_v
    assert np.abs(out["jacobian"].cpu().numpy() - blob["jacobian"]).max() <= tol_v * 5
    assert np.allclose(hm.sum((2, 3)), 1.0, atol=1e-4)


def test_source_cache_and_cuda_graph_replay_equal_eager(dev):
    """Opt-in encoder reuse for a repeated source tensor and CUDA-graph replay return the eager results."""
    from eamm_b200.graph import GraphedGenerator
    gen, cfg = generator("tiny", dev)
    gen.precision = "fp32"
    src, kpd, kps = synth.make_inputs(1, cfg, size=64, seed=60)
    frames = [synth.make_inputs(1, cfg, size=64, seed=61 + i)[1] for i in range(4)]
    s = src.to(dev)
    ks = to_dev(kps, dev)
    eager = [gen(s, kp_driving=to_dev(kd, dev), kp_source=ks)["prediction"].clone() for kd in frames]
    gen.cache_source = True
    try:
        cached = [gen(s, kp_driving=to_dev(kd, dev), kp_source=ks)["prediction"].clone() for kd in frames]
        s.mul_(1.0)                      # in-place write bumps the version counter -> the cache must miss
        again = gen(s, kp_driving=to_dev(frames[0], dev), kp_source=ks)["prediction"].clone()
    finally:
        gen.cache_source = False
    for a, b in zip(eager, cached):
        assert torch.equal(a, b)
    assert torch.equal(again, eager[0])
    graphed = GraphedGenerator(gen, s, to_dev(frames[0], dev), ks)
    for kd, want in zip(frames, eager):
        got = graphed(s, to_dev(kd, dev), ks)["prediction"]
        torch.cuda.synchronize()
        assert torch.equal(got, want)


# ------------------------------------------------------------------ per-clip keypoint glue
@pytest.mark.parametrize("name", ["kp_glue_emo_t12", "kp_glue_plain_t40"])
def test_kp_clip_glue_matches_reference_golden(dev, name):
    from eamm_b200 import clip
    blob = np.load(os.path.join(GOLD, name + ".npz"))
    T, with_emo = [int(v) for v in blob["meta"]]
    drv, emo, src, init = synth.make_clip_inputs(T=T)
    out = clip.smooth_and_normalize(to_dev(drv, dev), to_dev(src, dev), to_dev(init, dev),
                                    emo_driving_all=to_dev(emo, dev) if with_emo else None,
                                    relative=True, scale=float(blob["scale"][0]))
    torch.cuda.synchronize()
    assert np.abs(out["value"].cpu().numpy() - blob["value"]).max() <= 2e-6
    assert np.abs(out["jacobian"].cpu().numpy() - blob["jacobian"]).max() <= 5e-6
    assert abs(clip.movement_scale(src, init) - float(blob["scale"][0])) < 1e-6
    # the result is a valid batched kp_driving for the generator
    gen, cfg = generator("full", dev)
    gen.precision = "fp32"
    img = synth.make_inputs(1, cfg, size=256, seed=1)[0].to(dev)
    res = gen(img.expand(T, -1, -1, -1), kp_driving=out, kp_source={k: v.expand(T, *v.shape[1:]) for k, v in to_dev(src, dev).items()})
    assert res["prediction"].shape == (T, 3, 256, 256) and torch.isfinite(res["prediction"]).all()


@pytest.mark.parametrize("precision", ["fp32_simt", "fp32"])
def test_u8_frame_output_matches_img_as_ubyte(dev, precision):
    """optional uint8 NHWC frames from the final conv's epilogue."""
    from oracle import eamm_oracle as oracle
    gen, cfg = generator("full", dev)
    gen.precision = precision
    src, kpd, kps = synth.make_inputs(2, cfg, size=256, seed=70)
    gen.emit_u8 = True
    try:
        out = gen(src.to(dev), kp_driving=to_dev(kpd, dev), kp_source=to_dev(kps, dev))
    finally:
        gen.emit_u8 = False
    u8 = out["prediction_u8"].cpu()
    assert u8.dtype == torch.uint8 and u8.shape == (2, 256, 256, 3)
    assert torch.equal(u8, oracle.frames_u8(out["prediction"].cpu()))          # exact w.r.t. our own fp32 prediction
    want = oracle.frames_u8(oracle.generator_forward(synth.make_state_dict(cfg, seed=0), cfg, src, kpd, kps)["prediction"])
    assert (u8.int() - want.int()).abs().max() <= 1                            # at most one grey level vs the reference


def test_whole_clip_chain_matches_oracle_chain(dev):
    """Detector heads -> clip glue -> generator (shared source, batched over T) -> uint8 frames, i.e. the widened
    path of demo.py:206-281 for one clip, against the same chain evaluated with the CPU oracles."""
    from eamm_b200 import clip
    from eamm_b200.config import get_kp_config
    from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
    from oracle import eamm_oracle as oracle, kp_glue
    T = 6
    cfg = get_config("full")
    kcfg, acfg = get_kp_config("full"), get_kp_config("full", audio=True)
    sd, ksd, asd = synth.make_state_dict(cfg, seed=0), synth.make_kp_state_dict(kcfg, seed=2), synth.make_kp_state_dict(acfg, seed=3)
    src = synth.make_inputs(1, cfg, size=256, seed=80)[0]
    fmap = synth.make_kp_inputs(acfg, T, 256, True, seed=81) * 0.3           # stand-in for AT_net2's deco_out[:, t]
    # ---- oracle chain (CPU)
    o_src = oracle.kp_detector_forward(ksd, kcfg, src)
    o_drv = oracle.kp_detector_a_forward(asd, acfg, fmap)
    o_init = {k: o_drv[k][:1] for k in ("value", "jacobian")}
    scale = clip.movement_scale(o_src, o_init)
    nv, nj = kp_glue.clip_glue(o_drv["value"], o_drv["jacobian"], None, None, o_src, o_init, movement_scale=scale)
    o_out = oracle.generator_forward(sd, cfg, src.expand(T, -1, -1, -1).contiguous(), {"value": nv, "jacobian": nj},
                                     {k: o_src[k].expand(T, *o_src[k].shape[1:]).contiguous() for k in ("value", "jacobian")})
    # ---- CUDA chain
    gen, _ = generator("full", dev)
    gen.precision = "fp32"
    det = KPDetector(**kcfg).eval(); det.load_state_dict(ksd); det = det.to(dev); det.precision = "fp32"
    det_a = KPDetector_a(**acfg).eval(); det_a.load_state_dict(asd); det_a = det_a.to(dev); det_a.precision = "fp32"
    s = src.to(dev)
    k_src = det(s)
    k_drv = det_a(fmap.to(dev))
    k_init = {k: k_drv[k][:1] for k in ("value", "jacobian")}
    k_norm = clip.smooth_and_normalize(k_drv, k_src, k_init, relative=True, scale=clip.movement_scale(k_src, k_init))
    gen.emit_u8 = True
    try:
        out = gen(s.expand(T, -1, -1, -1), kp_driving=k_norm,
                  kp_source={k: k_src[k].expand(T, *k_src[k].shape[1:]) for k in ("value", "jacobian")})
    finally:
        gen.emit_u8 = False
    torch.cuda.synchronize()
    assert (k_norm["value"].cpu() - nv).abs().max() <= 5e-4
    assert (out["prediction"].cpu() - o_out["prediction"]).abs().max() <= 2e-3     # detector error enters through the flow
    assert (out["prediction_u8"].cpu().int() - oracle.frames_u8(o_out["prediction"]).int()).abs().max() <= 2


# ------------------------------------------------------------------ AT_net2 (BASELINE config 5)
_AT = {}


def at_net(dev):
    from eamm_b200.modules.util import AT_net2
    if "m" not in _AT:
        m = AT_net2().eval()
        m.load_state_dict(synth.make_at_state_dict(), strict=True)
        _AT["m"] = m.to(dev)
    return _AT["m"]


@pytest.mark.parametrize("name", ["at_b2_t3", "at_b1_t6"])
def test_at_net2_matches_reference_golden(dev, name):
    """MFCC conv encoder + pose MLP + image DownBlocks -> LSTM (cluster kernel) -> ConvTranspose stack, fp32.
    Tolerance 1e-4 max-abs on outputs of magnitude ~1 (fp32 sums in a different order than ATen's)."""
    blob = np.load(os.path.join(GOLD, name + ".npz"))
    B, T = [int(v) for v in blob["meta"]]
    img, mfcc, pose = synth.make_at_inputs(B, T)
    m = at_net(dev)
    out = m(img.to(dev), mfcc.to(dev), pose.to(dev), "cnn", 1.6)
    torch.cuda.synchronize()
    assert out.shape == (B, T, 35, 64, 64)
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got[..., ::4, ::4] - blob["out"]).max() <= 1e-4
    s = np.array([got.astype(np.float64).sum(), np.abs(got.astype(np.float64)).sum()])
    assert abs(s[1] - blob["sum_out"][1]) <= 2e-6 * blob["sum_out"][1]
    lstm = m._eng.ws[(B, T, 256, 256)].h[2].view(B, T, 256).cpu().numpy()
    assert np.abs(lstm - blob["lstm_out"]).max() <= 2e-5


def test_cabi_act_copy_moves_regions_between_formats_and_rezeroes_padding(dev):
    """eamm_act_copy (no reference counterpart: the padded-map plumbing of AT_net2's tensor-core MFCC convs)."""
    from eamm_b200 import _lib as L
    from eamm_b200.engine import ActBuf, current_stream_ptr
    lib = L.load()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 28, 12, 64, generator=g).to(dev)
    src = ActBuf(3, 28, 12, 64, "f32", dev)
    src.store_float(x)
    pad = ActBuf(3, 32, 16, 64, "bf16x2", dev)
    a, b = src.act(), pad.act()
    L.check(lib.eamm_act_copy(C.byref(a), C.byref(b), 28, 12, 0, current_stream_ptr()), "act_copy")
    got = pad.to_float()                                               # NCHW fp32
    assert (got[:, :, :28, :12] - x.permute(0, 3, 1, 2)).abs().max() <= 2e-5
    assert got[:, :, 28:].abs().max() == 0 and got[:, :, :, 12:].abs().max() == 0
    pad.t.fill_(1.0)                                                   # dirty everything, then restore the zero padding
    L.check(lib.eamm_act_copy(C.byref(b), C.byref(b), 26, 5, 1, current_stream_ptr()), "act_copy")
    t = pad.to_float()
    assert t[:, :, 26:].abs().max() == 0 and t[:, :, :, 5:].abs().max() == 0 and (t[:, :, :26, :5] != 0).all()
    back = ActBuf(3, 28, 12, 64, "f32", dev)
    pad.store_float(F.pad(x, (0, 0, 0, 4, 0, 4)))
    c = back.act()
    L.check(lib.eamm_act_copy(C.byref(b), C.byref(c), 28, 12, 0, current_stream_ptr()), "act_copy")
    torch.cuda.synchronize()
    assert (back.to_float() - x.permute(0, 3, 1, 2)).abs().max() <= 2e-5
    assert lib.eamm_act_copy(C.byref(a), C.byref(b), 40, 12, 0, current_stream_ptr()) == -1        # EAMM_ERR_ARG: region > source


def test_at_net2_long_clip_is_causal_and_batch_consistent(dev):
    """Size-independent properties at a clip length the oracle would take minutes for: frame t depends only on windows
    <= t (LSTM causality), and sequences of a batch do not interact."""
    T = 96
    img, mfcc, pose = synth.make_at_inputs(2, T, seed=9)
    m = at_net(dev)
    full = m(img.to(dev), mfcc.to(dev), pose.to(dev), "cnn", 1.6)
    head = m(img.to(dev), mfcc[:, :40].to(dev), pose[:, :40].to(dev), "cnn", 1.6)
    one = m(img[1:].to(dev), mfcc[1:].to(dev), pose[1:].to(dev), "cnn", 1.6)
    torch.cuda.synchronize()
    assert torch.isfinite(full).all()
    assert torch.equal(full[:, :40], head)
    assert torch.equal(full[1:], one)
    assert float((full[:, -1] - full[:, -2]).abs().mean()) > 1e-3


def test_audio_to_frames_chain_psnr(dev):
    """BASELINE.json config 5 in miniature: MFCC windows -> AT_net2 -> KPDetector_a per frame -> clip glue -> generator
    (demo.py:345, :206-281) against the same chain run with the CPU oracles.  Stated tolerance: PSNR >= 60 dB on
    [0,1] frames and max-abs 2e-3 (keypoint softmax at temperature 0.1 amplifies the fp32 rounding of the stages before)."""
    from eamm_b200 import clip
    from eamm_b200.config import get_kp_config
    from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
    from oracle import eamm_oracle as oracle, kp_glue
    T = 6
    cfg = get_config("full")
    kcfg, acfg = get_kp_config("full"), get_kp_config("full", audio=True)
    sd, ksd, asd = synth.make_state_dict(cfg, seed=0), synth.make_kp_state_dict(kcfg, seed=2), synth.make_kp_state_dict(acfg, seed=3)
    atsd = synth.make_at_state_dict()
    img, mfcc, pose = synth.make_at_inputs(1, T, seed=21)
    # ---- oracle chain (CPU)
    o_deco = oracle.at_net2_forward(atsd, img, mfcc, pose, 1.6)
    o_src = oracle.kp_detector_forward(ksd, kcfg, img)
    o_drv = oracle.kp_detector_a_forward(asd, acfg, o_deco[0])
    o_init = {k: o_drv[k][:1] for k in ("value", "jacobian")}
    scale = clip.movement_scale(o_src, o_init)
    nv, nj = kp_glue.clip_glue(o_drv["value"], o_drv["jacobian"], None, None, o_src, o_init, movement_scale=scale)
    o_out = oracle.generator_forward(sd, cfg, img.expand(T, -1, -1, -1).contiguous(), {"value": nv, "jacobian": nj},
                                     {k: o_src[k].expand(T, *o_src[k].shape[1:]).contiguous() for k in ("value", "jacobian")})
    # ---- CUDA chain
    gen, _ = generator("full", dev)
    gen.precision = "fp32"
    det = KPDetector(**kcfg).eval(); det.load_state_dict(ksd); det = det.to(dev); det.precision = "fp32_simt"
    det_a = KPDetector_a(**acfg).eval(); det_a.load_state_dict(asd); det_a = det_a.to(dev); det_a.precision = "fp32_simt"
    s = img.to(dev)
    deco = at_net(dev)(s, mfcc.to(dev), pose.to(dev), "cnn", 1.6)
    k_src = det(s)
    k_drv = det_a(deco[0])
    k_init = {k: k_drv[k][:1] for k in ("value", "jacobian")}
    k_norm = clip.smooth_and_normalize(k_drv, k_src, k_init, relative=True, scale=clip.movement_scale(k_src, k_init))
    out = gen(s.expand(T, -1, -1, -1), kp_driving=k_norm,
              kp_source={k: k_src[k].expand(T, *k_src[k].shape[1:]) for k in ("value", "jacobian")})
    torch.cuda.synchronize()
    assert (deco.cpu() - o_deco).abs().max() <= 1e-4
    assert (k_norm["value"].cpu() - nv).abs().max() <= 5e-4
    err = out["prediction"].cpu() - o_out["prediction"]
    psnr = -10.0 * torch.log10((err ** 2).mean())
    assert float(psnr) >= 60.0, float(psnr)
    assert err.abs().max() <= 2e-3
    # the clip actually moves: driven frames differ from each other
    assert float((o_out["prediction"][0] - o_out["prediction"][-1]).abs().mean()) > 1e-4
    # the one-call clip API (chunked over T, uint8 frames) gives the same frames
    det.precision = det_a.precision = "fp32_simt"
    frames = clip.animate_audio_clip(at_net(dev), det, det_a, gen, s, mfcc.to(dev), pose.to(dev), 1.6, chunk=4)
    assert frames.shape == (T, 256, 256, 3) and frames.dtype == torch.uint8
    assert (frames.cpu().int() - oracle.frames_u8(o_out["prediction"]).int()).abs().max() <= 2


def test_config4_real_mfcc_clip_300_frames_psnr(dev):
    """BASELINE.json configs[4]: the demo.py path on the reference's LRW sample (real MFCC windows [:, :, 1:] tiled to
    300 frames, demo.py:318-346) -> AT_net2 -> KPDetector_a -> One-Euro + normalize_kp -> generator, one call, uint8
    frames; against the frames the reference's own modules rendered (tests/golden/clip_lrw_t300.npz).  Stated
    tolerance: per-frame PSNR >= 50 dB on the uint8 frames, keypoints within 1e-3."""
    from test_oracle_golden import lrw_clip_case
    from eamm_b200 import clip
    from eamm_b200.config import get_kp_config
    from eamm_b200.modules.keypoint_detector import KPDetector, KPDetector_a
    blob, T, img, mfcc, pose = lrw_clip_case()
    kcfg, acfg = get_kp_config("full"), get_kp_config("full", audio=True)
    gen, _ = generator("full", dev)
    gen.precision = "fp32"
    det = KPDetector(**kcfg).eval(); det.load_state_dict(synth.make_kp_state_dict(kcfg, seed=2)); det = det.to(dev)
    det_a = KPDetector_a(**acfg).eval(); det_a.load_state_dict(synth.make_kp_state_dict(acfg, seed=3)); det_a = det_a.to(dev)
    det.precision = det_a.precision = "fp32_simt"
    s = img.to(dev)
    frames = clip.animate_audio_clip(at_net(dev), det, det_a, gen, s, mfcc.to(dev), pose.to(dev), 1.6, chunk=60)
    torch.cuda.synchronize()
    assert frames.shape == (T, 256, 256, 3) and frames.dtype == torch.uint8
    got = frames[:, ::8, ::8].cpu().float()
    want = torch.from_numpy(blob["frames_u8_s8"]).float()
    mse = ((got - want) ** 2).mean(dim=(1, 2, 3))
    worst = float(10 * torch.log10(255.0 ** 2 / mse.clamp_min(1e-12)).min())
    print("configs[4] clip: worst per-frame PSNR %.1f dB, max |diff| %d grey levels, %.2f%% of sampled bytes identical"
          % (worst, int((got - want).abs().max()), 100 * float((got == want).float().mean())))
    assert worst >= 50.0
    # the keypoints that drove it (recomputed: animate_audio_clip returns frames only)
    deco = at_net(dev)(s, mfcc.to(dev), pose.to(dev), "cnn", 1.6)
    k_src, k_drv = det(s), det_a(deco[0])
    k_init = {k: k_drv[k][:1] for k in ("value", "jacobian")}
    k_norm = clip.smooth_and_normalize(k_drv, k_src, k_init, relative=True, scale=clip.movement_scale(k_src, k_init))
    kerr = float(np.abs(k_norm["value"].cpu().numpy() - blob["kp_value"]).max())
    print("configs[4] clip: normalised keypoints max-abs %.3e (tol 1.0e-03)" % kerr)
    assert kerr <= 1e-3


def test_config2_batch256_fp16_matches_oracle_on_sampled_frames(dev):
    """BASELINE.json configs[2] at its own size: 256 frames in ONE call, one-pass fp16 convs / fp32 warp (kept last in the
    file: the largest case).  The reference is batch-invariant to 6e-8, so the CPU oracle runs on a sample of
    the frames (first / last, both sides of the 32- and 128-frame boundaries) instead of 25 s on all of them; the structural
    properties are checked on all 256.  Tolerance = SURVEY 8(c)'s reduced-precision bound (1e-2 and PSNR >= 50 dB)."""
    from oracle import eamm_oracle as oracle
    cfg = get_config("full")
    B = 256
    src, kpd, kps = synth.make_inputs(B, cfg, size=256, seed=23)
    got = run_ours("full", dev, "fp16", src, kpd, kps)
    for k in ("prediction", "deformed", "mask", "occlusion_map", "sparse_deformed"):
        assert got[k].shape[0] == B and bool(torch.isfinite(got[k]).all()), k
    assert torch.allclose(got["mask"].sum(1), torch.ones(B, 64, 64), atol=1e-5)
    assert got["occlusion_map"].min() >= 0 and got["occlusion_map"].max() <= 1
    assert got["prediction"].min() >= 0 and got["prediction"].max() <= 1
    idx = [0, 1, 31, 32, 127, 128, 200, 255]
    pick = lambda d: {k: v[idx] for k, v in d.items()}
    want = oracle.generator_forward(synth.make_state_dict(cfg, seed=0), cfg, src[idx], pick(kpd), pick(kps))
    for k, tol in TOL["fp16"].items():
        err = (got[k][idx] - want[k]).abs().max().item()
        print("configs[2] B=256 fp16 mode vs oracle (8 sampled frames): %-16s max-abs %.3e (tol %.1e)" % (k, err, tol))
        assert err <= tol, (k, err)
    p = psnr(got["prediction"][idx], want["prediction"])
    print("configs[2] B=256 fp16 mode prediction PSNR %.1f dB" % p)
    assert p >= 50.0
