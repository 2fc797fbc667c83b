"""Known-answer tests that pin the ORACLE (CPU, no GPU needed).

The reference ships no golden vectors for this path. Besides the pass-by-pass comparison with the reference's
own shaders (tests/test_reference_shaders.py), the oracle is anchored on
the analytic properties the shaders imply — each test cites the shader lines that imply the expected value.
"""
import ctypes as C
import math

import numpy as np
import pytest

from diligentfx_b200 import capi, synth
from oracle import oracle_py as op


@pytest.fixture(scope="module")
def L(built):
    return op.lib()


# ---------------------------------------------------------------------------------------------------------------------
# ToneMapping (config 1 of BASELINE.json: Uncharted2 on a 256x256 synthetic HDR buffer, scalar host loop)
# ---------------------------------------------------------------------------------------------------------------------
def _u2(x):
    A, B, Cc, D, E, F = 0.15, 0.50, 0.10, 0.20, 0.02, 0.30
    return ((x * (A * x + Cc * B) + D * E) / (x * (A * x + B) + D * F)) - E / F


def test_tonemap_uncharted2_closed_form_256x256(L):
    """ToneMapping.fxh:8-19, :125-132 — curr = U2(2*scale*c) / U2(whitePoint), scale = middleGray/aveLogLum."""
    rng = np.random.default_rng(1)
    c = np.exp2(rng.uniform(-8.0, 6.0, (256, 256, 3))).astype(np.float32)
    a = capi.ToneMapAttribs.default()
    got = op.tone_map(a, 0.3, c)
    want = _u2(2.0 * (0.18 / 0.3) * c.astype(np.float64)) / _u2(3.0)
    assert np.abs(got - want).max() < 2e-6
    # U2(inf) = 1 - E/F, so the operator saturates at (1 - E/F) / U2(whitePoint)
    assert got.min() >= 0.0 and got.max() < (1.0 - 0.02 / 0.30) / _u2(3.0)


def test_tonemap_kats(L):
    a = capi.ToneMapAttribs.default()
    # Uncharted2Tonemap(0) = D*E/(D*F) - E/F = 0  (ToneMapping.fxh:18)
    assert abs(op.tone_map(a, 0.3, np.zeros((1, 3), np.float32))).max() < 1e-7
    # white-point normalisation: 2*scale*c == whitePoint  ->  1.0
    c = np.full((1, 3), 3.0 / (2.0 * 0.18 / 0.3), np.float32)
    assert np.abs(op.tone_map(a, 0.3, c) - 1.0).max() < 1e-6
    # negative input is clamped first (f3Color = max(f3Color, 0), :95)
    assert abs(op.tone_map(a, 0.3, np.full((1, 3), -5.0, np.float32))).max() < 1e-7
    # mode NONE returns max(c, 0)
    a.iToneMappingMode = 0
    assert np.allclose(op.tone_map(a, 0.3, np.array([[0.25, 2.0, -1.0]], np.float32)), [[0.25, 2.0, 0.0]])
    # Reinhard: lum-preserving form L/(1+L) * c/lum (:109-123)
    a.iToneMappingMode = 2
    c = np.array([[0.5, 0.5, 0.5]], np.float32)
    lum = 0.5 * (0.212671 + 0.715160 + 0.072169)
    ls = lum * 0.18 / 0.3
    assert np.allclose(op.tone_map(a, 0.3, c), (ls / (1 + ls)) * (c / lum), atol=1e-6)
    # every operator is finite and non-negative on a positive grid
    grid = np.exp2(np.linspace(-8, 6, 64, dtype=np.float32))[:, None].repeat(3, 1)
    for mode in range(12):
        a = capi.ToneMapAttribs.default()
        a.iToneMappingMode = mode
        out = op.tone_map(a, 0.3, grid)
        assert np.isfinite(out).all(), mode
        if mode not in (8, 9):  # the AgX polynomial fit dips slightly below 0 near black
            assert out.min() >= -1e-6, mode


# ---------------------------------------------------------------------------------------------------------------------
# integer / scalar helpers
# ---------------------------------------------------------------------------------------------------------------------
def test_pcg_hash_kat(L):
    """PostFX_Common.fxh:20-25, evaluated with Python big-int arithmetic mod 2^32."""
    def pcg(s):
        st = (s * 747796405 + 2891336453) & 0xFFFFFFFF
        w = (((st >> ((st >> 28) + 4)) ^ st) * 277803737) & 0xFFFFFFFF
        return ((w >> 22) ^ w) & 0xFFFFFFFF
    for s in (0, 1, 2, 12345, 0xFFFFFFFF, 0x80000000):
        assert L.orc_pcg_hash(C.c_uint32(s)) == pcg(s)


def test_bayer4x4_is_the_bayer_matrix(L):
    """PostFX_Common.fxh:57-65: the packed constants decode to the 4x4 ordered-dither matrix; frame index rotates it mod 16."""
    m = np.array([[L.orc_bayer4x4(x, y, 0) * 16 for x in range(4)] for y in range(4)])
    assert sorted(m.ravel().tolist()) == list(range(16))
    assert m[0, 0] == 0
    # classic Bayer: every 2x2 block holds one value from each quarter of the range
    for by in (0, 2):
        for bx in (0, 2):
            assert sorted((m[by:by + 2, bx:bx + 2].ravel() // 4).astype(int).tolist()) == [0, 1, 2, 3]
    for f in (1, 5, 17):
        m2 = np.array([[L.orc_bayer4x4(x, y, f) * 16 for x in range(4)] for y in range(4)])
        assert np.array_equal(m2, (m + f) % 16)
    assert L.orc_bayer4x4(5, 7, 3) == L.orc_bayer4x4(1, 3, 3)  # wraps with period 4


def test_halton_and_jitter(L):
    """TemporalAntiAliasing.cpp:43-78: Halton(2,3), 16-frame cycle, jitter within +-1 pixel in NDC."""
    assert L.orc_halton(2, 1) == 0.5 and L.orc_halton(2, 2) == 0.25 and L.orc_halton(2, 3) == 0.75
    assert abs(L.orc_halton(3, 1) - 1 / 3) < 1e-7 and abs(L.orc_halton(3, 2) - 2 / 3) < 1e-7 and abs(L.orc_halton(3, 4) - 4 / 9) < 1e-7
    out = (C.c_float * 2)()
    for f in range(40):
        L.orc_taa_jitter(f, 1920, 1080, out)
        assert abs(out[0]) <= 1.0 / 1920 * 1.0001 and abs(out[1]) <= 1.0 / 1080 * 1.0001
        jx, jy = synth.taa_jitter(f, 1920, 1080)
        assert out[0] == pytest.approx(jx, abs=1e-9) and out[1] == pytest.approx(jy, abs=1e-9)
        out2 = (C.c_float * 2)()
        L.orc_taa_jitter(f + 16, 1920, 1080, out2)
        assert (out[0], out[1]) == (out2[0], out2[1])


def test_fast_acos_error(L):
    """SSAO_ComputeAmbientOcclusion.fx:47-53: the sqrt-based fit stays within ~0.02 rad of acos on [-1, 1]."""
    xs = np.linspace(-1, 1, 4001)
    err = max(abs(L.orc_fast_acos(C.c_float(x)) - math.acos(x)) for x in xs)
    assert err < 0.02
    assert abs(L.orc_fast_acos(C.c_float(1.0))) < 1e-6 and abs(L.orc_fast_acos(C.c_float(-1.0)) - math.pi) < 1e-6


def test_depth_camera_z_roundtrip(L):
    """ShaderUtilities.fxh:5-39 with the D3D LH projection of SURVEY.md Appendix B.1: near -> 0, far -> 1, inverse pair."""
    cam = synth.make_camera(0, 640, 360, use_jitter=False).attribs
    P = cam.mProj
    assert L.orc_camera_z_to_depth(C.c_float(synth.NEAR), C.byref(P)) == pytest.approx(0.0, abs=1e-6)
    assert L.orc_camera_z_to_depth(C.c_float(synth.FAR), C.byref(P)) == pytest.approx(1.0, abs=1e-6)
    for z in (0.1, 0.5, 1.0, 7.3, 50.0, 100.0):
        d = L.orc_camera_z_to_depth(C.c_float(z), C.byref(P))
        assert L.orc_depth_to_camera_z(C.c_float(d), C.byref(P)) == pytest.approx(z, rel=2e-3)


def test_bloom_mip_count(L):
    """Bloom.cpp:152-156 / SURVEY.md Appendix C: 8 levels at 4K, 7 at 1080p, 9 at 8K for Radius 0.75."""
    assert L.orc_bloom_mip_count(1920, 1080, C.c_float(0.75)) == 8
    assert L.orc_bloom_mip_count(960, 540, C.c_float(0.75)) == 7
    assert L.orc_bloom_mip_count(3840, 2160, C.c_float(0.75)) == 9


# ---------------------------------------------------------------------------------------------------------------------
# pass-level analytic cases
# ---------------------------------------------------------------------------------------------------------------------
def _plane_scene(w, h, z=5.0):
    """A wall facing the camera at view depth z, zero motion, static camera."""
    cam = synth.make_camera(0, w, h, use_jitter=False)
    # un-pitched camera for this case: identity view
    a = cam.attribs
    ident = np.eye(4, dtype=np.float32)
    for name in ("mView", "mViewInv"):
        for r in range(4):
            for c in range(4):
                getattr(a, name).m[r][c] = float(ident[r, c])
    P = np.array([[a.mProj.m[r][c] for c in range(4)] for r in range(4)], np.float64)
    for name, M in (("mViewProj", P), ("mViewProjInv", np.linalg.inv(P))):
        for r in range(4):
            for c in range(4):
                getattr(a, name).m[r][c] = float(M[r, c])
    a.f4Position[:] = [0, 0, 0, 1]
    d = (z * P[2, 2] + P[3, 2]) / z
    depth = np.full((h, w), d, np.float32)
    normal = np.zeros((h, w, 4), np.float32)
    normal[..., 2] = -1.0
    color = np.full((h, w, 4), 0.5, np.float32)
    material = np.zeros((h, w, 4), np.float32)
    motion = np.zeros((h, w, 2), np.float32)
    return dict(depth=depth, normal=normal, color=color, material=material, motion=motion, prev_depth=depth.copy(), curr_camera=a, prev_camera=a, frame=0)


def test_ssao_plane_and_background(built):
    """SSAO_ComputeAmbientOcclusion.fx:188-226: on a plane facing the camera both horizons stay at +-90 deg, so the
    cosine-weighted arc integral gives visibility 1 (up to the FastACos fit); :139-140 + clear :982 -> background AO = 1."""
    w, h = 96, 64
    fr = _plane_scene(w, h)
    fr["depth"][:8, :] = 1.0  # sky band
    o = op.Oracle(w, h, threads=2)
    o.set_inputs(fr)
    o.run("blue_noise"), o.run("ssao_prefilter"), o.run("ssao_ao")
    ao = o.get("ssao_occ")
    assert np.all(ao[:8] == 1.0)
    inner = ao[24:-8, 16:-16]
    assert np.abs(inner - 1.0).max() < 0.02, (inner.min(), inner.max())
    for algo in (1, 2):  # HBAO, VBAO also see an unoccluded hemisphere
        a = capi.SSAOAttribs.default()
        a.Algorithm = algo
        o.set_ssao(a)
        o.run("ssao_ao")
        inner = o.get("ssao_occ")[24:-8, 16:-16]
        assert np.abs(inner - 1.0).max() < 0.05, (algo, inner.min(), inner.max())


def test_hiz_is_block_min_with_odd_edges(built):
    """SSR_ComputeHierarchicalDepthBuffer.fx:52-70: mip k texel = min over its 2^k block; odd source sizes fold the extra
    row / column into the last texel."""
    rng = np.random.default_rng(5)
    w, h = 77, 45
    d = rng.uniform(0.2, 0.99, (h, w)).astype(np.float32)
    o = op.Oracle(w, h, threads=1)
    o.set("depth", d)
    o.run("ssr_hiz")
    prev = d
    for k in range(1, 7):
        m = o.get(f"ssr_hiz.{k}")
        ph, pw = prev.shape
        assert m.shape == (max(h >> k, 1), max(w >> k, 1))
        want = np.empty_like(m)
        for y in range(m.shape[0]):
            for x in range(m.shape[1]):
                x1 = 2 * x + 2 + (1 if pw & 1 else 0)
                y1 = 2 * y + 2 + (1 if ph & 1 else 0)
                want[y, x] = prev[2 * y:min(y1, ph), 2 * x:min(x1, pw)].min()
        assert np.array_equal(m, want), k
        prev = m
    # global property: the top level is <= everything it covers
    assert o.get("ssr_hiz.6").min() >= d.min()


def test_bloom_below_threshold_is_identity(built):
    """Bloom_ComputePrefilteredTexture.fx:24-35: constant colour below Threshold - Knee contributes nothing, so the
    composite returns the source colour."""
    w, h = 128, 72
    o = op.Oracle(w, h, threads=2)
    c = np.zeros((h, w, 4), np.float32)
    c[..., :3] = [0.3, 0.5, 0.2]
    o.set("bloom_in", c)
    o.run("bloom")
    assert np.abs(o.get("bloom_down0")).max() == 0.0
    assert np.allclose(o.get("bloom_out")[..., :3], c[..., :3], atol=1e-7)
    # above the threshold the pyramid is non-zero and energy spreads but the centre stays close to src + Intensity*prefiltered
    c[..., :3] = 4.0
    o.set("bloom_in", c)
    o.run("bloom")
    out = o.get("bloom_out")[h // 2, w // 2, :3]
    assert np.all(out > 4.0)


def test_taa_constant_colour_fixed_point(built):
    """TAA_…fx:229-261: zero motion + constant colour is a fixed point; first frame resets with alpha 0.5 (:235-236),
    afterwards alpha' = min(stability, 1/(2-alpha)) (:224-227)."""
    w, h = 64, 48
    fr = _plane_scene(w, h)
    fr["color"][..., :3] = [0.7, 1.5, 0.1]
    o = op.Oracle(w, h, threads=2)
    alpha = None
    for f in range(4):
        fr["frame"] = f
        fr["curr_camera"].uiFrameIndex = f
        o.set_inputs(fr)
        o.frame(op.STAGE_POSTFX | op.STAGE_TAA)
        acc = o.get(f"taa_accum{f & 1}")
        assert np.allclose(acc[..., :3], fr["color"][..., :3], rtol=2e-5, atol=2e-6), f
        want = 0.5 if f == 0 else min(0.9375, 1.0 / (2.0 - alpha))
        assert np.allclose(acc[..., 3], want, atol=1e-6), (f, acc[..., 3].mean(), want)
        alpha = want


def test_closest_motion_border_reads_zero(built):
    """ComputeClosestMotion.fx:34-35 + SURVEY.md Appendix B.2: unclamped Loads return 0 out of bounds, so border pixels see
    a 'closest' depth of 0 off-screen and fetch the (zero) off-screen motion."""
    w, h = 32, 16
    d = np.full((h, w), 0.5, np.float32)
    m = np.ones((h, w, 2), np.float32)
    o = op.Oracle(w, h, threads=1)
    o.set("depth", d), o.set("motion", m)
    o.run("closest_motion")
    cm = o.get("closest_motion")
    assert np.all(cm[1:-1, 1:-1] == 1.0)
    assert np.all(cm[0, :] == 0.0) and np.all(cm[-1, :] == 0.0) and np.all(cm[:, 0] == 0.0) and np.all(cm[:, -1] == 0.0)


def test_blue_noise_ranges_and_determinism(built):
    """ComputeBlueNoiseTexture.fx:20-79: values in [0,1), frame-dependent, reproducible; XY uses Sobol^tile of dims 0,1."""
    o = op.Oracle(128, 128, threads=1)
    o.set_frame_index(0)
    o.run("blue_noise")
    xy0, zw0 = o.get("bn_xy"), o.get("bn_zw")
    o.set_frame_index(1)
    o.run("blue_noise")
    xy1 = o.get("bn_xy")
    for a in (xy0, zw0, xy1):
        assert a.shape == (128, 128, 2) and a.min() >= 0.0 and a.max() < 1.0
    assert not np.array_equal(xy0, xy1)
    blob = np.frombuffer(open(op.TABLES, "rb").read(), np.uint8)
    sob, tile = blob[:256], blob[256:]
    x, y = 37, 101
    v = (float(sob[0] ^ tile[(x + y * 128) * 8 + 0]) + 0.5) / 256.0
    assert xy0[y, x, 0] == pytest.approx((v + 0.5) % 1.0, abs=1e-6)
    # roughly uniform
    assert abs(xy0.mean() - 0.5) < 0.02 and abs(zw0.mean() - 0.5) < 0.02


def test_full_chain_runs_and_is_deterministic(built, seq_even):
    o1, o2 = op.Oracle(256, 144, threads=4), op.Oracle(256, 144, threads=1)
    for fr in seq_even[:2]:
        o1.set_inputs(fr), o2.set_inputs(fr)
        o1.frame(), o2.frame()
    a, b = o1.get("ldr"), o2.get("ldr")
    assert np.array_equal(a, b)  # thread count must not change results
    # fp32 storage: no UNORM clamp, Uncharted2 saturates at ~1.99 -> sRGB 1.35
    assert np.isfinite(a).all() and 0.0 <= a[..., :3].min() and a[..., :3].max() <= 1.36
