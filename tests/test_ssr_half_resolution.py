"""ScreenSpaceReflection::FEATURE_FLAG_HALF_RESOLUTION (ScreenSpaceReflection.hpp:64-76): S3 downsampled mask
(SSR_ComputeDownsampledStencilMask.fx:13-61), S4 one ray per 2x2 block into width/2 x height/2 targets
(SSR_ComputeIntersection.fx:283-288, pattern PostFX_Common.fxh:45-55), S5 gathering from the half-size targets
(SSR_ComputeSpatialReconstruction.fx:153-157); S6 / S7 unchanged.

CPU: known answers of the oracle for the 4x4 offset pattern and the mask rule. GPU: S3 bit-exact, S4 / S5 and the whole chain
against the oracle through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

from helpers import Dev, assert_close, psnr, reinhard, rows
from diligentfx_b200 import capi, synth

W, H, FRAMES = 160, 96, 3


def _oracle(w=W, h=H, threads=4):
    from oracle import oracle_py as op
    return op.Oracle(w, h, threads=threads)


def test_half_resolution_offset_pattern():
    # PostFX_Common.fxh:47-51 prints the matrix the packed constant encodes; index = ((x & 3) << 3) + ((y & 3) << 1)
    packed = 1320229860
    got = [[(packed >> (((x & 3) << 3) + ((y & 3) << 1))) & 3 for y in range(4)] for x in range(4)]
    assert got == [[0, 1, 2, 3], [3, 2, 1, 0], [1, 0, 3, 2], [2, 3, 0, 1]]
    assert all(sorted(r) == [0, 1, 2, 3] for r in got) and all(sorted(c) == [0, 1, 2, 3] for c in zip(*got))  # a Latin square


def test_oracle_downsampled_mask_rule(built):
    o = _oracle(4, 4, 1)
    depth = np.full((4, 4), 0.5, np.float32)
    depth[0:2, 0:2] = 1.0                       # block (0,0): all background -> fails
    depth[0, 2] = 1.0                           # block (1,0): one background texel, closest depth 0.5 -> passes
    rough = np.zeros((4, 4), np.float32)
    rough[2, 0] = 0.9                           # block (0,1): max roughness above the 0.2 threshold -> fails
    o.set("depth", depth), o.set("ssr_roughness", rough)
    o.run("ssr_downsample_mask")
    assert o.get("ssr_mask_half").tolist() == [[0.0, 1.0], [0.0, 1.0]]


@pytest.fixture(scope="module")
def half(built):
    seq = synth.generate_sequence(W, H, FRAMES)
    o = _oracle()
    o.set_ssr(capi.SSRAttribs.default(), capi.SSR_FLAG_HALF_RESOLUTION)
    for fr in seq:
        o.set_inputs(fr)
        o.frame()
    return seq, o


def test_oracle_half_resolution_frame_structure(half):
    _, o = half
    assert o.get("ssr_mask_half").shape == (H // 2, W // 2) and o.get("ssr_radiance").shape == (H // 2, W // 2, 4)
    assert o.get("ssr_resolved_rad").shape == (H, W, 4) and o.get("ssr_out").shape == (H, W, 4)
    out = o.get("ssr_out")
    assert np.isfinite(out).all() and out[..., 3].max() > 0.1      # some confident reflections survive


@pytest.mark.gpu
def test_cuda_half_resolution_passes(half):
    seq, o = half
    fr = seq[-1]
    d = Dev()
    cams = d.cameras(fr["curr_camera"], fr["prev_camera"])
    a = capi.SSRAttribs.default()
    hw, hh = W // 2, H // 2
    depth, rough = d.up(fr["depth"]), d.up(o.get("ssr_roughness"))
    # S3: a selection, bit-exact
    mh = d.empty(hh, hw, dtype=d.torch.uint8, fill=7)
    capi.check(d.lib.dfx_pass_ssr_downsample_mask(None, C.byref(a), C.byref(d.plane(rough)), C.byref(d.plane(depth)), C.byref(d.plane(mh)), rows(hh)), "S3")
    d.sync()
    assert np.array_equal(d.host(mh), o.get("ssr_mask_half"))
    # S4 on the oracle's pyramid / mask
    hiz = [depth] + [d.up(o.get(f"ssr_hiz.{i}")) for i in range(1, 7)]
    rad, rdir = d.empty(hh, hw, 4, fill=-1.0), d.empty(hh, hw, 4, fill=-1.0)
    capi.check(d.lib.dfx_pass_ssr_intersect(None, cams, C.byref(a), capi.SSR_FLAG_HALF_RESOLUTION, C.byref(d.plane(d.up(fr["color"]))),
                                            C.byref(d.plane(d.up(fr["normal"]))), C.byref(d.plane(rough)), C.byref(d.plane(d.mask(o.get("ssr_mask_half")))),
                                            C.byref(d.plane(d.up(o.get("bn_xy")))), C.byref(d.pyr(hiz)), None, C.byref(d.plane(rad)), C.byref(d.plane(rdir)),
                                            rows(hh)), "S4")
    d.sync()
    # The Hi-Z march is a chaotic integer walk (an ulp in the ray set-up can move a hit by a cell): compare statistically,
    # exactly like the full-resolution test in test_parity_gpu.py
    grad, wrad, gdir, wdir = d.host(rad), o.get("ssr_radiance"), d.host(rdir), o.get("ssr_raydir")
    m = o.get("ssr_mask_half") > 0
    assert np.all(grad[~m] == 0.0) and np.all(gdir[~m] == 0.0), "masked-out pixels keep the clear value"
    ghit, whit = grad[..., 3] > 0, wrad[..., 3] > 0
    both = m & ghit & whit
    dlen = np.abs(np.linalg.norm(gdir[..., :3], axis=-1) - np.linalg.norm(wdir[..., :3], axis=-1))
    same = both & (dlen < 1e-2 * (1.0 + np.linalg.norm(wdir[..., :3], axis=-1)))
    assert (ghit == whit)[m].mean() >= 0.99
    assert same.sum() >= 0.98 * both.sum()
    assert_close("half-res ssr pdf", gdir[..., 3], wdir[..., 3], tol=1e-3 * max(1.0, float(np.abs(wdir[..., 3]).max())), max_outliers=5e-3, mask=m)
    assert_close("half-res ssr radiance (same hit)", grad, wrad, tol=5e-3, max_outliers=5e-3, hdr=True, mask=same)
    # S5 gathering from the oracle's half-size targets
    res, var, dep = d.empty(H, W, 4, fill=0.0), d.empty(H, W, fill=0.0), d.empty(H, W, fill=0.0)
    capi.check(d.lib.dfx_pass_ssr_spatial(None, cams, C.byref(a), C.byref(d.plane(rough)), C.byref(d.plane(d.mask(o.get("ssr_mask")))),
                                          C.byref(d.plane(d.up(fr["normal"]))), C.byref(d.plane(depth)), C.byref(d.plane(d.up(o.get("ssr_raydir")))),
                                          C.byref(d.plane(d.up(o.get("ssr_radiance")))), C.byref(d.plane(res)), C.byref(d.plane(var)), C.byref(d.plane(dep)),
                                          rows(H)), "S5")
    d.sync()
    m = o.get("ssr_mask") > 0
    assert_close("half-res resolved radiance", d.host(res), o.get("ssr_resolved_rad"), tol=5e-3, max_outliers=5e-3, hdr=True, mask=m)
    # mismatched sizes are refused
    bad = d.empty(hh, hw + 1, 4)
    assert d.lib.dfx_pass_ssr_spatial(None, cams, C.byref(a), C.byref(d.plane(rough)), C.byref(d.plane(d.mask(o.get("ssr_mask")))),
                                      C.byref(d.plane(d.up(fr["normal"]))), C.byref(d.plane(depth)), C.byref(d.plane(bad)), C.byref(d.plane(bad)),
                                      C.byref(d.plane(res)), C.byref(d.plane(var)), C.byref(d.plane(dep)), rows(H)) == capi.DFX_ERR_INVALID_ARG


@pytest.mark.gpu
def test_cuda_half_resolution_chain(half):
    from diligentfx_b200.chain import ChainConfig, PostProcessChain
    seq, o = half
    chain = PostProcessChain(W, H, ChainConfig(ssr_flags=capi.SSR_FLAG_HALF_RESOLUTION))
    for fr in seq:
        ldr = chain.run_frame(fr).cpu().numpy()
    assert chain.fetch("ssr", 3).shape == (H // 2, W // 2, 4)
    assert np.array_equal(chain.fetch("ssr", 20), o.get("ssr_mask_half"))
    assert psnr(reinhard(chain.fetch("ssr", 0)), reinhard(o.get("ssr_out"))) >= 38.0
    assert psnr(np.clip(ldr[..., :3], 0, 1), np.clip(o.get("ldr")[..., :3], 0, 1)) >= 49.0
    chain.close()
