"""DepthOfField: the eleven passes of PostProcess/DepthOfField/src/DepthOfField.cpp:292-331
(Shaders/PostProcess/DepthOfField/private/DOF_*.fx), between TAA and Bloom in the reference chain.

CPU: known answers of the oracle (thin-lens CoC formula, kernel point counts, pass-through of an in-focus frame, energy of the
Gaussian). GPU: every pass against the oracle through the C-ABI on the oracle's own inputs, then the effect object inside
the chain."""
import ctypes as C

import numpy as np
import pytest

from helpers import Dev, assert_close, psnr, reinhard, rows
from diligentfx_b200 import capi, synth

W, H, FRAMES = 160, 96, 3
FLAGS = capi.DOF_FLAG_TEMPORAL_SMOOTHING | capi.DOF_FLAG_KARIS_INVERSE


def _oracle(w=W, h=H, threads=4):
    from oracle import oracle_py as op
    return op.Oracle(w, h, threads=threads)


def _lens(fr, focus=6.0, fstop=1.4):
    for k in ("curr_camera", "prev_camera"):
        fr[k].fFocusDistance, fr[k].fFStop, fr[k].fFocalLength, fr[k].fSensorWidth = focus, fstop, 50.0, 36.0
    return fr


def _attribs():
    a = capi.DOFAttribs.default()
    a.MaxCircleOfConfusion = 0.02
    return a


def test_oracle_coc_formula(built):
    """CoC = clamp(1000 * K * (z - F) / z / (sensor * MaxCoC), -1, 1), K = f^2 / (N * (F - f)), f in metres (…CircleOfConfusion.fx:24-39)."""
    from oracle import oracle_py as op
    fr = _lens(synth.generate_sequence(32, 16, 1)[0], focus=5.0, fstop=2.0)
    o = _oracle(32, 16, 1)
    a = _attribs()
    o.set_dof(a, 0)
    o.set_inputs(fr)
    o.frame(op.STAGE_POSTFX | op.STAGE_DOF)
    got, depth = o.get("dof_coc"), fr["depth"]
    proj = fr["curr_camera"].mProj
    z = np.array([[o.L.orc_depth_to_camera_z(C.c_float(float(d)), C.byref(proj)) for d in row] for row in depth], np.float64)
    f, N, F = 0.05, 2.0, 5.0
    want = np.clip(1000.0 * (f * f / (N * (F - f))) * (z - F) / np.maximum(z, 1e-4) / (36.0 * a.MaxCircleOfConfusion), -1, 1)
    assert np.allclose(got, want, atol=2e-5)
    assert got.min() < 0 < got.max()                          # the scene straddles the focus plane


def test_kernel_point_counts():
    # ComputeSampleCount (DOF_Common.fx:4-7) == points generated (DepthOfField.cpp:49-73): 1 + density * rings * (rings - 1) / 2
    for rings in range(2, 6):
        for density in range(2, 8):
            gen = sum(max(density * i, 1) for i in range(rings))
            assert gen == 1 + density * ((rings - 1) * rings >> 1) <= 71


def test_oracle_in_focus_frame_passes_through(built):
    """With the whole scene at the focus distance every CoC is 0: both layer alphas are 0 and D11 returns the source."""
    from oracle import oracle_py as op
    fr = synth.generate_sequence(64, 40, 1)[0]
    fr["depth"][:] = fr["depth"][20, 32]                      # a flat wall
    z = _oracle(64, 40).L.orc_depth_to_camera_z(C.c_float(float(fr["depth"][0, 0])), C.byref(fr["curr_camera"].mProj))
    _lens(fr, focus=float(z))
    o = _oracle(64, 40)
    o.set_dof(_attribs(), 0)
    o.set_inputs(fr)
    o.frame(op.STAGE_POSTFX | op.STAGE_DOF)
    assert np.abs(o.get("dof_coc")).max() < 1e-3
    assert np.allclose(o.get("dof_out"), o.get("dof_in"), atol=1e-6)


@pytest.fixture(scope="module")
def ref(built):
    from oracle import oracle_py as op
    seq = [_lens(f) for f in synth.generate_sequence(W, H, FRAMES)]
    o = _oracle()
    o.set_dof(_attribs(), FLAGS)
    for fr in seq:
        o.set_inputs(fr)
        o.frame(op.STAGE_ALL | op.STAGE_DOF)
    return seq, o


def test_oracle_frame_structure(ref):
    _, o = ref
    assert o.get("dof_dilation3").shape == (H >> 3, W >> 3) and o.get("dof_pre0").shape == (H // 2, W // 2, 4)
    src, out = o.get("dof_in"), o.get("dof_out")
    assert np.isfinite(out).all() and np.array_equal(out[..., 3], src[..., 3])
    # defocus removes high frequencies: the Laplacian energy of the blurred frame is lower
    lap = lambda x: np.abs(4 * x[1:-1, 1:-1] - x[:-2, 1:-1] - x[2:, 1:-1] - x[1:-1, :-2] - x[1:-1, 2:]).mean()  # noqa: E731
    assert lap(reinhard(out[..., :3])) < 0.8 * lap(reinhard(src[..., :3]))
    assert np.isfinite(o.get("ldr")).all()


@pytest.mark.gpu
def test_cuda_dof_passes(ref):
    seq, o = ref
    fr = seq[-1]
    cur, prv = fr["frame"] & 1, (fr["frame"] + 1) & 1
    d, a = Dev(), _attribs()
    L = d.lib
    cams = d.cameras(fr["curr_camera"], fr["prev_camera"])
    P = lambda arr: C.byref(d.plane(d.up(arr)))  # noqa: E731
    hw, hh = W // 2, H // 2
    # D1
    coc = d.empty(H, W)
    capi.check(L.dfx_pass_dof_coc(None, cams, C.byref(a), P(fr["depth"]), C.byref(d.plane(coc)), rows(H)), "D1")
    assert_close("coc", d.host(coc), o.get("dof_coc"), tol=2e-5)
    # D2 on the oracle's planes (previous slot = history of frame - 1)
    tc = d.empty(H, W)
    capi.check(L.dfx_pass_dof_temporal_coc(None, cams, C.byref(a), P(o.get("dof_coc")), P(o.get(f"dof_coc_temporal{prv}")), P(o.get("closest_motion")),
                                           C.byref(d.plane(tc)), rows(H)), "D2")
    assert_close("temporal coc", d.host(tc), o.get(f"dof_coc_temporal{cur}"), tol=2e-5, max_outliers=1e-3)
    # D3, D4 x3: selections, bit-exact
    s0 = d.empty(H, W)
    capi.check(L.dfx_pass_dof_separated_coc(None, P(o.get(f"dof_coc_temporal{cur}")), C.byref(d.plane(s0)), rows(H)), "D3")
    d.sync()
    assert np.array_equal(d.host(s0), o.get("dof_dilation0"))
    for k in range(1, 3):                                      # level 3 is overwritten by the blur in the oracle's frame: check 1 and 2
        src = o.get(f"dof_dilation{k - 1}")
        out = d.empty(src.shape[0] >> 1, src.shape[1] >> 1)
        capi.check(L.dfx_pass_dof_dilation(None, P(src), C.byref(d.plane(out)), rows(out.shape[0])), "D4")
        d.sync()
        assert np.array_equal(d.host(out), o.get(f"dof_dilation{k}")), k
    # D4 (level 3) + D5 + D6 chained, against the blurred level 3
    l3, tmp = d.empty(H >> 3, W >> 3), d.empty(H >> 3, W >> 3)
    capi.check(L.dfx_pass_dof_dilation(None, P(o.get("dof_dilation2")), C.byref(d.plane(l3)), rows(H >> 3)), "D4.3")
    capi.check(L.dfx_pass_dof_blur_coc(None, C.byref(d.plane(l3)), 0, C.byref(d.plane(tmp)), rows(H >> 3)), "D5")
    capi.check(L.dfx_pass_dof_blur_coc(None, C.byref(d.plane(tmp)), 1, C.byref(d.plane(l3)), rows(H >> 3)), "D6")
    assert_close("blurred dilation", d.host(l3), o.get("dof_dilation3"), tol=1e-6)
    assert L.dfx_pass_dof_blur_coc(None, C.byref(d.plane(l3)), 1, C.byref(d.plane(l3)), rows(H >> 3)) == capi.DFX_ERR_INVALID_ARG  # not in-place
    # D7
    coc_t = o.get(f"dof_coc_temporal{cur}")
    fg, bg = d.empty(hh, hw, 4), d.empty(hh, hw, 4)
    capi.check(L.dfx_pass_dof_prefilter(None, P(o.get("dof_in")), P(coc_t), P(o.get("dof_dilation3")), C.byref(d.plane(fg)), C.byref(d.plane(bg)), rows(hh)), "D7")
    # the oracle's dof_pre* hold the state after D9; recompute D7 .. D11 here step by step from the GPU's own planes instead, and
    # compare the planes the oracle still has at the end of its frame: bokeh0/1 (after D10) and the output
    b0, b1 = d.empty(hh, hw, 4), d.empty(hh, hw, 4)
    capi.check(L.dfx_pass_dof_bokeh(None, cams, C.byref(a), FLAGS, 0, C.byref(d.plane(fg)), C.byref(d.plane(bg)), P(o.get("dof_in")), C.byref(d.plane(b0)),
                                    C.byref(d.plane(b1)), rows(hh)), "D8")
    capi.check(L.dfx_pass_dof_bokeh(None, cams, C.byref(a), FLAGS, 1, C.byref(d.plane(b0)), C.byref(d.plane(b1)), None, C.byref(d.plane(fg)),
                                    C.byref(d.plane(bg)), rows(hh)), "D9")
    assert_close("foreground after the flood fill", d.host(fg), o.get("dof_pre0"), tol=1e-4, max_outliers=2e-3, hdr=True)
    assert_close("background after the flood fill", d.host(bg), o.get("dof_pre1"), tol=1e-4, max_outliers=2e-3, hdr=True)
    capi.check(L.dfx_pass_dof_postfilter(None, C.byref(d.plane(fg)), C.byref(d.plane(bg)), C.byref(d.plane(b0)), C.byref(d.plane(b1)), rows(hh)), "D10")
    assert_close("near layer", d.host(b0), o.get("dof_bokeh0"), tol=1e-4, max_outliers=2e-3, hdr=True)
    assert_close("far layer", d.host(b1), o.get("dof_bokeh1"), tol=1e-4, max_outliers=2e-3, hdr=True)
    out = d.empty(H, W, 4)
    capi.check(L.dfx_pass_dof_combine(None, C.byref(a), P(o.get("dof_in")), C.byref(d.plane(b0)), C.byref(d.plane(b1)), C.byref(d.plane(out)), rows(H)), "D11")
    assert_close("depth of field output", d.host(out), o.get("dof_out"), tol=1e-4, max_outliers=2e-3, min_psnr=70.0, hdr=True)


@pytest.mark.gpu
def test_cuda_dof_in_the_chain(ref):
    from diligentfx_b200.chain import ChainConfig, PostProcessChain
    seq, o = ref
    chain = PostProcessChain(W, H, ChainConfig(dof=_attribs(), dof_flags=FLAGS))
    for fr in seq:
        ldr = chain.run_frame(fr).cpu().numpy()
    assert psnr(chain.fetch("dof", 1), o.get("dof_coc")) >= 80.0
    assert psnr(reinhard(chain.fetch("dof", 0)), reinhard(o.get("dof_out"))) >= 45.0
    assert psnr(np.clip(ldr[..., :3], 0, 1), np.clip(o.get("ldr")[..., :3], 0, 1)) >= 45.0
    # and it is a different picture from the chain without it
    plain = PostProcessChain(W, H)
    for fr in seq:
        ldr0 = plain.run_frame(fr).cpu().numpy()
    assert psnr(np.clip(ldr[..., :3], 0, 1), np.clip(ldr0[..., :3], 0, 1)) < 40.0
    chain.close(), plain.close()
