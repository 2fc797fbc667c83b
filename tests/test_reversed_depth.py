"""PostFXContext::FEATURE_FLAG_REVERSED_DEPTH (depth near = 1, far = 0; the reference compiles *_OPTION_INVERTED_DEPTH shader
variants: ComputeClosestMotion.fx:5-9,36-40; SSAO_Common.fxh:6-23; SSR_Common.fxh:6-12,48-55; SSR_ComputeIntersection.fx:109-124).

The same synthetic scene is rendered with a forward and with a reversed depth buffer (`synth.reverse_depth_frame`:
depth' = 1 - depth, projection with m22' = 1 - m22, m32' = -m32). The two encodings describe the same geometry, so
the chain must produce (nearly) the same frame from either — a size-independent property that needs no second
implementation — and the CUDA chain must match the oracle in reversed mode as closely as it does in forward mode."""
import numpy as np
import pytest

W, H, FRAMES = 160, 96, 3


def psnr(a, b, peak=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 200.0 if mse == 0 else 10 * np.log10(peak * peak / mse)


def rein(x):
    x = np.maximum(x, 0)
    return x / (1 + x)


@pytest.fixture(scope="module")
def frames():
    from diligentfx_b200 import synth
    fwd = synth.generate_sequence(W, H, FRAMES)
    return fwd, [synth.reverse_depth_frame(f) for f in fwd]


def run_oracle(seq, reversed_depth: bool) -> dict:
    from oracle import oracle_py as op
    o = op.Oracle(W, H, threads=4)
    o.set_reversed_depth(reversed_depth)
    try:
        for fr in seq:
            o.set_inputs(fr)
            o.frame()
        last = (FRAMES - 1) & 1
        return {"ldr": o.get("ldr"), "ssao": o.get("ssao_out"), "ssr": o.get("ssr_out"), "closest": o.get("closest_motion"),
                "taa": o.get(f"taa_accum{last}")}
    finally:
        o.set_reversed_depth(False)


def test_oracle_reversed_equals_forward(built, frames):
    fwd, rev = frames
    a, b = run_oracle(fwd, False), run_oracle(rev, True)
    # closest motion: the 3x3 search picks the same neighbour (the ordering of 1 - d is the exact mirror of the ordering of d)
    # except on the frame border, where out-of-bounds loads return 0 = nearest in forward and farthest in reversed mode
    assert np.array_equal(a["closest"][1:-1, 1:-1], b["closest"][1:-1, 1:-1])
    assert psnr(a["ssao"], b["ssao"]) >= 45.0
    assert psnr(rein(a["ssr"]), rein(b["ssr"])) >= 35.0
    assert psnr(np.clip(a["ldr"][..., :3], 0, 1), np.clip(b["ldr"][..., :3], 0, 1)) >= 40.0
    # and the flag matters: reversed frames processed in forward mode are a different (wrong) picture
    c = run_oracle(rev, False)
    assert psnr(np.clip(a["ldr"][..., :3], 0, 1), np.clip(c["ldr"][..., :3], 0, 1)) < 35.0


@pytest.mark.gpu
def test_cuda_reversed_depth_chain(built, frames):
    from diligentfx_b200 import capi
    from diligentfx_b200.chain import ChainConfig, PostProcessChain
    fwd, rev = frames
    want = run_oracle(rev, True)
    cr = PostProcessChain(W, H, ChainConfig(postfx_flags=capi.POSTFX_FLAG_REVERSED_DEPTH))
    cf = PostProcessChain(W, H)
    for f, r in zip(fwd, rev):
        ldr_r = cr.run_frame(r).cpu().numpy()
        ldr_f = cf.run_frame(f).cpu().numpy()
    # parity with the oracle in reversed mode: the floors of the forward chain test (tests/test_parity_gpu.py)
    assert psnr(np.clip(ldr_r[..., :3], 0, 1), np.clip(want["ldr"][..., :3], 0, 1)) >= 49.0
    assert psnr(cr.fetch("ssao", 0), want["ssao"]) >= 50.0
    assert psnr(rein(cr.fetch("ssr", 0)), rein(want["ssr"])) >= 40.0
    assert psnr(rein(cr.fetch("taa", 0)), rein(want["taa"])) >= 40.0
    assert np.array_equal(cr.fetch("postfx", 4), want["closest"])  # closest motion is a selection: bit-exact
    # the property: both encodings give the same picture
    assert psnr(np.clip(ldr_r[..., :3], 0, 1), np.clip(ldr_f[..., :3], 0, 1)) >= 40.0
    cr.close(), cf.close()


@pytest.mark.gpu
def test_cuda_reversed_hiz_is_max_pyramid(built):
    """S1 with the plane flag: every level is the 2x2 (+ odd row / column) MAXIMUM, far plane 0 (SSR_Common.fxh:6-12)."""
    import ctypes as C

    import torch
    from diligentfx_b200 import capi
    L = capi.load()
    rng = np.random.default_rng(11)
    w, h = 75, 51  # odd sizes exercise the extra row / column rule
    d = rng.random((h, w), dtype=np.float32)
    lv = [torch.from_numpy(d).cuda()] + [torch.full((max(h >> i, 1), max(w >> i, 1)), -1.0, device="cuda") for i in range(1, 7)]
    pyr = capi.pyramid_of(lv, flags=capi.PLANE_FLAG_REVERSED_DEPTH)
    capi.check(L.dfx_pass_ssr_hiz(None, C.byref(pyr), capi.Rows(0, h)), "hiz")
    torch.cuda.synchronize()
    prev = d
    for i in range(1, 7):
        ph, pw = prev.shape
        hh, ww = max(ph >> 1, 1), max(pw >> 1, 1)
        want = np.zeros((hh, ww), np.float32)
        for y in range(hh):
            for x in range(ww):
                ys = [min(2 * y + k, ph - 1) for k in range(3 if ph & 1 else 2)]
                xs = [min(2 * x + k, pw - 1) for k in range(3 if pw & 1 else 2)]
                want[y, x] = max(0.0, max(prev[yy, xx] for yy in ys for xx in xs))
        assert np.array_equal(lv[i].cpu().numpy(), want), f"level {i}"
        prev = want


@pytest.mark.gpu
def test_temporal_upscaling_runs_bloom_at_output_resolution(built):
    """PostFXContext::FEATURE_FLAG_TEMPORAL_UPSCALING: Bloom takes FrameDesc.OutputWidth x OutputHeight instead of the render
    resolution (Bloom.cpp:84-85). Checked through the effect-level objects against the oracle's Bloom at that size."""
    import ctypes as C

    import torch
    from diligentfx_b200 import capi
    from diligentfx_b200.capi import BloomAttribs, BloomRenderAttribs, FrameDesc, Plane
    from oracle import oracle_py as op
    L = capi.load()
    w, h, ow, oh = 96, 64, 144, 96  # 1.5x upscaling
    rng = np.random.default_rng(3)
    src = np.exp2(rng.uniform(-3, 5, (oh, ow, 4))).astype(np.float32)
    o = op.Oracle(ow, oh, threads=2)
    o.set_bloom(BloomAttribs.default())
    o.set("bloom_in", src)
    o.run("bloom")
    postfx, bloom = C.c_void_p(), C.c_void_p()
    capi.check(L.dfx_postfx_create(C.byref(postfx))), capi.check(L.dfx_bloom_create(C.byref(bloom)))
    try:
        assert L.dfx_postfx_prepare(postfx, C.byref(FrameDesc(0, w, h, 0, 0)), 4) == capi.DFX_ERR_INVALID_ARG  # needs the output size
        capi.check(L.dfx_postfx_prepare(postfx, C.byref(FrameDesc(0, w, h, ow, oh)), 4), "postfx prepare")
        capi.check(L.dfx_bloom_prepare(bloom, postfx, 0), "bloom prepare")
        color = torch.from_numpy(src).cuda()
        pc, a = capi.plane_of(color), BloomAttribs.default()
        capi.check(L.dfx_bloom_execute(bloom, C.byref(BloomRenderAttribs(None, postfx, C.pointer(pc), C.pointer(a)))), "bloom execute")
        out = Plane()
        capi.check(L.dfx_bloom_get_plane(bloom, 0, C.byref(out)))
        assert (out.width, out.height) == (ow, oh)
        from diligentfx_b200.chain import download_plane
        got = download_plane(out)
        assert psnr(rein(got), rein(o.get("bloom_out"))) >= 90.0
    finally:
        L.dfx_bloom_destroy(bloom), L.dfx_postfx_destroy(postfx)
