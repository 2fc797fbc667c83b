"""Pins the oracle against the reference ITSELF: every pass of the oracle is compared with the reference's own HLSL pixel
shader for that pass, compiled for the CPU by oracle/refshader (sources read in place from /root/reference, never copied) and
run on the same inputs.

Bar: BIT-EXACT. The shader runner and the oracle state the HLSL intrinsics and the fixed-function sampler the same way
(fp32, no contraction, 1/256-texel bilinear snap), so any difference is a difference between the oracle's restatement and
the reference's shader text. The one tolerated pass is the first bokeh gather of DepthOfField, whose 142 bilinear taps per
pixel are aimed with a texture coordinate the shader derives from the interpolated NDC position (0.5 + 0.5 * ndc) while
the oracle uses (x + 0.5) / W: the last-bit difference moves a few taps across a 1/256-texel snap boundary.

Runs wherever oracle/_ref/librefshaders.so exists: here it is (re)built from /root/reference; on a box without the
reference the prebuilt library that travelled with the snapshot is used; with neither the module is skipped and
tests/test_reference_shader_golden.py still checks the committed outputs of these shaders.
"""
import numpy as np
import pytest

from diligentfx_b200 import capi, synth
from oracle.refshader import refsh

from oracle.refshader.driver import Variant, compare_frame, make_oracle

W, H = 157, 89                                   # odd on both axes: the odd-size branches of every mip chain run
TOLERATED = {"D8 dof_bokeh_first.near": (1e-3, 0.02), "D8 dof_bokeh_first.far": (1e-3, 0.02)}   # (max abs, max fraction of texels that differ)


@pytest.fixture(scope="module")
def shaders(built):
    try:
        refsh.build()
    except Exception as e:                       # pragma: no cover - only where the reference is mounted but does not compile
        pytest.fail(f"the reference shaders no longer compile for the CPU: {e}")
    if not refsh.available():
        pytest.skip("oracle/_ref/librefshaders.so is absent and /root/reference is not mounted")
    return True


def _run(v: Variant, frames=None, warm: int = 2):
    seq = frames if frames is not None else synth.generate_sequence(W, H, warm + 1)
    h, w = seq[0]["depth"].shape
    o = make_oracle(w, h, v)
    o.set_reversed_depth(v.reversed_depth)
    try:
        for fr in seq[:warm]:
            o.set_inputs(fr)
            o.frame(v.stages())
        return compare_frame(o, seq[warm], v)
    finally:
        o.set_reversed_depth(False)


def _check(res):
    bad = []
    for label, (got, want) in res.items():
        assert got.shape == want.shape, label
        if label in TOLERATED:
            tol, frac = TOLERATED[label]
            d = np.abs(got.astype(np.float64) - want)
            if d.max() > tol or (d > 0).mean() > frac:
                bad.append(f"{label}: max abs {d.max():.3e}, differing texels {(d > 0).mean():.3%}")
        elif not np.array_equal(np.ascontiguousarray(got).view(np.uint32), np.ascontiguousarray(want).view(np.uint32)):  # bit patterns: NaN == NaN
            d = np.abs(got.astype(np.float64) - want)
            bad.append(f"{label}: NOT bit-exact (max abs {d.max():.3e}, {(d > 0).mean():.3%} of values differ)")
    assert not bad, "oracle differs from the reference's shaders:\n  " + "\n  ".join(bad)
    return len(res)


def test_default_chain_bit_exact(shaders):
    """The benchmarked configuration: every pass of PostFX, SSR, SSAO, TAA, Bloom and ToneMap, third frame of a sequence."""
    n = _check(_run(Variant()))
    assert n >= 45


def test_reversed_depth_bit_exact(shaders):
    seq = [synth.reverse_depth_frame(f) for f in synth.generate_sequence(W, H, 3)]
    _check(_run(Variant(reversed_depth=True), seq))


def test_half_resolution_bit_exact(shaders):
    _check(_run(Variant(ssr_flags=capi.SSR_FLAG_HALF_RESOLUTION, ssao_flags=capi.SSAO_FLAG_HALF_RESOLUTION)))


def test_ssr_previous_frame_and_half_precision_depth_bit_exact(shaders):
    _check(_run(Variant(ssr_flags=1, ssao_flags=1)))


@pytest.mark.parametrize("algorithm", [1, 2], ids=["hbao", "vbao"])
def test_ssao_algorithms_bit_exact(shaders, algorithm):
    res = _run(Variant(ssao_algorithm=algorithm))
    _check({k: v for k, v in res.items() if k.startswith("A")})


@pytest.mark.parametrize("flags", [0, 1, 4, 5, 6, 7], ids=lambda f: f"taa_flags_{f}")
def test_taa_variants_bit_exact(shaders, flags):
    res = _run(Variant(taa_flags=flags))
    _check({k: v for k, v in res.items() if k.startswith("T")})


@pytest.mark.parametrize("mode", [1, 2, 3, 5, 6, 7, 8, 9, 10, 11])
def test_tone_map_operators_bit_exact(shaders, mode):
    res = _run(Variant(tonemap_mode=mode, to_srgb=bool(mode & 1)), warm=0)
    _check({k: v for k, v in res.items() if k.startswith("M")})


def test_depth_of_field_bit_exact(shaders):
    a = capi.DOFAttribs.default()
    a.MaxCircleOfConfusion = 0.02
    frames = []
    for f in synth.generate_sequence(W, H, 3):
        g = dict(f)
        for k in ("curr_camera", "prev_camera"):
            c = capi.CameraAttribs.from_buffer_copy(bytes(f[k]))
            c.fFocusDistance, c.fFStop = 6.0, 1.4
            g[k] = c
        frames.append(g)
    for flags in (capi.DOF_FLAG_TEMPORAL_SMOOTHING | capi.DOF_FLAG_KARIS_INVERSE, 0):
        res = _run(Variant(dof=True, dof_flags=flags, dof_attribs=a), frames)
        assert any(k.startswith("D11") for k in res)
        _check({k: v for k, v in res.items() if k[0] in "DB"})


@pytest.mark.parametrize("size", [(34, 18), (33, 17)], ids=["even", "odd"])
def test_bilateral_quad_derivatives_everywhere(shaders, size):
    """SSR_ComputeBilateralCleanup.fx takes ddx / ddy of the camera-space depth. With the filter branch live on EVERY pixel
    (the sequence tests reach it only where the variance is high) the oracle's 2x2 differences equal the shader's quad
    derivatives bit for bit - including, on odd-sized targets, the last column / row, whose quad partner lies outside the
    target: Direct3D runs it as a helper lane whose Load returns 0. (The oracle and dfx_ssr.cu used to clamp that partner onto
    the pixel itself; this comparison is what showed the difference, and both were changed to the Load-returns-0 rule.)"""
    from oracle import oracle_py as op
    w, h = size
    fr = synth.generate_sequence(w, h, 1)[0]
    rng = np.random.default_rng(1)
    ssr = capi.SSRAttribs.default()
    o = op.Oracle(w, h)
    o.set_ssr(ssr, 0)
    o.set_inputs(fr)
    depth = rng.uniform(0.3, 0.9, (h, w)).astype(np.float32)
    normal = np.zeros((h, w, 4), np.float32)
    normal[..., 2] = 1.0
    o.set("depth", depth), o.set("normal", normal), o.set("material", np.full((h, w, 4), 0.2, np.float32))
    o.run("ssr_mask")
    rough, mask = o.get("ssr_roughness"), o.get("ssr_mask")
    assert mask.all() and rough.min() >= 0.125                            # every pixel traced, full filter radius
    ci = fr["frame"] & 1
    rad, var = rng.uniform(0, 4, (h, w, 4)).astype(np.float32), np.ones((h, w), np.float32)
    o.set(f"ssr_radhist{ci}", rad), o.set(f"ssr_varhist{ci}", var)
    o.run("ssr_bilateral")
    want = o.get("ssr_out")
    assert np.abs(want - rad).max() > 0.1                                  # the filter really ran
    got = np.zeros_like(want)
    refsh.run("ssr_bilateral", [depth, normal, rough, rad, var], [got], cbs=[fr["curr_camera"], ssr], mask=np.ones((h, w), np.uint8))
    assert np.array_equal(got, want), f"{(got != want).any(axis=2).sum()} pixels differ"


@pytest.mark.parametrize("trial", range(6))
def test_random_configurations_bit_exact(shaders, trial):
    """Random frame size, scene seed, feature flags, tone-map operator and attribute values (seeded): every pass of the third frame
    must still equal its shader bit for bit."""
    rng = np.random.default_rng(2024 + trial)
    w, h = int(rng.integers(40, 220)), int(rng.integers(30, 130))
    v = Variant(ssr_flags=int(rng.integers(0, 4)), ssao_flags=int(rng.integers(0, 4)), ssao_algorithm=int(rng.integers(0, 3)), taa_flags=int(rng.integers(0, 8)),
                tonemap_mode=int(rng.integers(1, 12)), to_srgb=bool(rng.integers(0, 2)), dof=bool(rng.integers(0, 2)), dof_flags=int(rng.integers(0, 4)))
    if v.ssr_flags == 3:
        v.ssr_flags = 2                                   # the manifest builds previous-frame and half-resolution intersect separately
    if v.ssao_algorithm:
        v.ssao_flags = 0                                  # HBAO / VBAO are built for the full-resolution, full-precision path
    if v.ssao_flags == 3:
        v.ssao_flags = 2
    v.ssr.MostDetailedMip, v.ssr.MaxTraversalIntersections = int(rng.integers(0, 3)), int(rng.integers(8, 200))
    v.ssr.RoughnessThreshold, v.ssr.GGXImportanceSampleBias = float(rng.uniform(0.05, 0.6)), float(rng.uniform(0, 1))
    v.ssr.SpatialReconstructionRadius, v.ssao.SpatialReconstructionRadius = float(rng.uniform(0, 6)), float(rng.uniform(0, 6))
    v.ssao.EffectRadius, v.ssao.TemporalStabilityFactor = float(rng.uniform(0.2, 3)), float(rng.uniform(0, 1))
    v.bloom.Threshold, v.bloom.Intensity, v.bloom.Radius = float(rng.uniform(0, 2)), float(rng.uniform(0, 1)), float(rng.uniform(0.5, 1.0))
    v.taa.TemporalStabilityFactor, v.dof_attribs.MaxCircleOfConfusion = float(rng.uniform(0, 1)), float(rng.uniform(0.005, 0.03))
    seq = synth.generate_sequence(w, h, 3, seed=int(rng.integers(1, 1000)))
    if v.dof:
        focus, fstop = float(rng.uniform(2, 12)), float(rng.uniform(1.2, 8))
        lens = []
        for f in seq:
            g = dict(f)
            for k in ("curr_camera", "prev_camera"):
                c = capi.CameraAttribs.from_buffer_copy(bytes(f[k]))
                c.fFocusDistance, c.fFStop = focus, fstop
                g[k] = c
            lens.append(g)
        seq = lens
    assert _check(_run(v, seq)) >= 40


def test_brdf_table_bit_exact(shaders):
    """PrecomputeBRDF.psh against oracle_compose_ibl.cpp's brdf_lut(), at two sample counts."""
    from oracle import oracle_py as op
    o = op.Oracle(W, H)
    for size, samples, name in ((64, 512, "brdf_lut"), (32, 64, "brdf_lut__64")):     # powers of two, like the reference's 512 x 512 table
        o.brdf_lut(size, samples)
        want = o.get("brdf_lut")
        got = np.zeros_like(want)
        refsh.run(name, [], [got])
        _check({f"BRDF table {size} x {size}, {samples} samples": (got, want)})


@pytest.mark.parametrize("size", [(128, 64), (157, 89)], ids=["128x64", "157x89"])
def test_full_compose(shaders, size):
    """Hydrogent's HnPostProcess.psh (SSR re-weighted by the split-sum BRDF and exchanged for the specular IBL, then SSAO)
    against the oracle's compose_ibl, on planes that exercise opacity 0, out-of-range roughness and metals.

    The shader builds the view direction from the interpolated NDC position of the pixel; the runner interpolates it as
    -1 + 2 (x + 0.5) / W, the oracle as ((x + 0.5) / W - 0.5) / 0.5. For power-of-two sizes both are exact and the pass is
    bit-exact; for other sizes the last bit of the NDC position differs and the result agrees to 1e-5 (relative to 1 + |c|)."""
    from oracle import oracle_py as op
    w, h = size
    fr = synth.generate_sequence(w, h, 1)[0]
    rng = np.random.default_rng(5)
    color = fr["color"].copy()
    color[..., 3] = rng.choice([0.0, 0.35, 1.0], (h, w), p=[0.1, 0.2, 0.7])
    mat = fr["material"].copy()
    mat[..., 1] = rng.choice([0.0, 0.5, 1.0], (h, w))
    mat[..., 0] = np.where(rng.random((h, w)) < 0.05, 1.5, mat[..., 0])
    base = np.concatenate([rng.uniform(0.02, 1.0, (h, w, 3)), np.ones((h, w, 1))], -1).astype(np.float32)
    ibl = np.concatenate([np.exp2(rng.uniform(-4, 2, (h, w, 3))), np.ones((h, w, 1))], -1).astype(np.float32)
    ssr = np.concatenate([np.exp2(rng.uniform(-4, 3, (h, w, 3))), rng.uniform(0, 1, (h, w, 1))], -1).astype(np.float32)
    ao = rng.uniform(0.2, 1.0, (h, w)).astype(np.float32)
    o = op.Oracle(w, h)
    o.set_inputs(fr)
    o.brdf_lut(64, 512)
    lut = o.get("brdf_lut")
    for k, v in dict(color=color, material=mat, base_color=base, specular_ibl=ibl, ssr_out=ssr, ssao_out=ao).items():
        o.set(k, v)
    bits = lambda f: int(np.float32(f).view(np.uint32))  # noqa: E731
    for ssr_scale, ssao_scale in ((0.8, 0.6), (1.0, 0.0), (0.0, 1.0)):
        o.set_compose_scales(ssr_scale, ssao_scale)
        o.run("compose_ibl")
        want = o.get("composed")
        got = np.zeros_like(want)
        refsh.run("compose_ibl", [color, ssr, ao, fr["normal"], ibl, mat, base, lut], [got], cbs=[fr["curr_camera"]], iparams=[bits(ssr_scale), bits(ssao_scale)])
        if w & (w - 1) == 0 and h & (h - 1) == 0:
            _check({f"compose (SSRScale {ssr_scale}, SSAOScale {ssao_scale})": (got, want)})
        else:
            rel = np.abs(got.astype(np.float64) - want) / (1.0 + np.abs(want))
            assert rel.max() < 1e-5, rel.max()


def test_constant_buffer_layouts_match_the_reference_structures(shaders):
    """Each harness memcpy's the C struct of include/dfx_b200.h into the structure the reference's .fxh declares and refuses
    (error 2) when the sizes differ; feeding a short struct must therefore fail."""
    import ctypes as C

    class Short(C.Structure):
        _fields_ = [("x", C.c_float * 3)]

    d = np.zeros((8, 8), np.float32)
    with pytest.raises(RuntimeError, match="failed with 2"):
        refsh.run("postfx_reprojected_depth", [d], [d.copy()], cbs=[Short(), Short()])
