"""GPU parity at the sizes BASELINE.json's configurations are quoted on (the tests of test_parity_gpu.py run at <= 333x187):

  config 2  1920x1080  SSAO A1-A8, every pass isolated on the oracle's own input planes (4th frame of a sequence);
  config 3  3840x2160  the whole chain, four consecutive frames, LDR >= 49 dB + per-stage PSNR floors;
  config 4  7680x4320  SSR S1-S7, every pass isolated on the oracle's planes (3rd frame, PostFX + SSR stages only).

Size-dependent paths these reach and the small tests do not: the 8-level Bloom pyramid with its five exact-2:1 levels, AO taps
at prefiltered-depth mips 3-4 (the pixel radius scales with the resolution), the 16-bit packing of Hi-Z level sizes, 32-bit texel
indexing of 530 MB planes, the 1/512-texel margin of the TAA Catmull-Rom taps at 3840 / 7680 columns.
The pass bodies are the ones of test_parity_gpu.py, called with a context of the larger size (same tolerances).
"""
import numpy as np
import pytest

import test_parity_gpu as P
from diligentfx_b200 import synth

pytestmark = pytest.mark.gpu


def _ctx(w, h, frames, stages=None):
    from oracle import oracle_py as op
    seq = synth.generate_sequence(w, h, frames)
    o = op.Oracle(w, h)
    for fr in seq:
        o.set_inputs(fr)
        o.frame(stages if stages is not None else op.STAGE_ALL)
    return dict(w=w, h=h, seq=seq, fr=seq[-1], o=o, f=seq[-1]["frame"])


@pytest.fixture(scope="module")
def ctx1080(built):
    return _ctx(1920, 1080, 4)


@pytest.fixture(scope="module")
def ctx4k(built):
    return _ctx(3840, 2160, 4)


@pytest.fixture(scope="module")
def ctx8k(built):
    from oracle import oracle_py as op
    return _ctx(7680, 4320, 3, op.STAGE_POSTFX | op.STAGE_SSR)


# ---- config 2: SSAO + bilateral blur at 1920x1080 ---------------------------------------------------------------------
def test_config2_ssao_prefilter_1080p(ctx1080):
    P.test_ssao_prefilter(ctx1080)


def test_config2_ssao_ambient_occlusion_1080p(ctx1080):
    P.test_ssao_ambient_occlusion(ctx1080, 0)


def test_config2_ssao_temporal_1080p(ctx1080):
    P.test_ssao_temporal(ctx1080)


def test_config2_ssao_convolute_resample_spatial_1080p(ctx1080):
    P.test_ssao_convolute_resample_spatial(ctx1080)


def test_config2_ssao_effect_1080p(ctx1080):
    """The effect object (all of A1-A8 in sequence, four frames) against the oracle's final AO plane."""
    from diligentfx_b200.chain import STAGE_POSTFX, STAGE_SSAO, ChainConfig, PostProcessChain
    from helpers import assert_close
    o, seq, h, w = ctx1080["o"], ctx1080["seq"], ctx1080["h"], ctx1080["w"]
    chain = PostProcessChain(w, h, ChainConfig(stages=STAGE_POSTFX | STAGE_SSAO))
    for fr in seq:
        chain.run_frame(fr)
    print(assert_close("ssao output 1080p", chain.fetch("ssao", 0), o.get("ssao_out"), tol=1e-2, max_outliers=2e-2, min_psnr=50.0))
    chain.close()


# ---- config 3: the whole chain at 3840x2160 ------------------------------------------------------------------------------
def test_config3_full_chain_4k(ctx4k):
    P.test_full_chain_four_frames(ctx4k)


def test_config3_bloom_passes_4k(ctx4k):
    P.test_bloom_passes(ctx4k)


def test_config3_taa_4k(ctx4k):
    P.test_taa(ctx4k, 2)


def test_config3_fused_chain_4k(ctx4k):
    """The benchmarked configuration (fused compose + TAA, fused composite + tone map, async compute) at 4K: LDR PSNR."""
    from diligentfx_b200.chain import PostProcessChain
    from helpers import psnr
    o, seq, h, w = ctx4k["o"], ctx4k["seq"], ctx4k["h"], ctx4k["w"]
    chain = PostProcessChain(w, h)
    for fr in seq:
        ldr = chain.run_frame(fr)
    got, want = ldr.cpu().numpy(), o.get("ldr")
    p = psnr(np.clip(got[..., :3], 0, 1), np.clip(want[..., :3], 0, 1))
    print(f"benchmarked chain, LDR PSNR after 4 frames at {w}x{h}: {p:.2f} dB")
    assert p >= 49.0
    chain.close()


# ---- config 4: SSR at 7680x4320 ----------------------------------------------------------------------------------------
def test_config4_ssr_hiz_and_mask_8k(ctx8k):
    P.test_ssr_hiz_and_mask(ctx8k)


def test_config4_ssr_intersect_8k(ctx8k):
    P.test_ssr_intersect(ctx8k, 0)


def test_config4_ssr_spatial_temporal_bilateral_8k(ctx8k):
    P.test_ssr_spatial_temporal_bilateral(ctx8k)


def test_config4_postfx_prepare_8k(ctx8k):
    P.test_postfx_prepare(ctx8k)
