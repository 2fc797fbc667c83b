"""cuemu — DEVELOPMENT TOOL (not collected by a plain `pytest tests/`). The whole chain on the HOST build of the kernels
(tools/cuemu), against the oracle, at degenerate frame sizes and with extreme attribute values; meant to be run under
AddressSanitizer, so that an indexing error at an edge nobody tests on the GPU shows up as a report, not as a wrong pixel:

    python tools/cuemu/build_emu.py --asan
    CUEMU_ASAN=1 LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \\
        python tests/cuemu_chain_sweep.py

End of round 1: every size from 17x9 up (incl. 64x1 and 1x64), every attribute setting and every feature-flag combination
below (reversed depth x half-resolution SSR x half-resolution / half-precision SSAO x DepthOfField, 97x55) agrees with the
oracle at >= 84 dB with no sanitizer report (VBAO, whose sector bits flip one at a time, at >= 62 dB); sizes whose Bloom pyramid
has fewer than two levels are refused with an error, as is a Bloom radius that leaves fewer than two levels. The sweep found
one limitation, since lifted: half-resolution SSAO refused a tightly pitched depth plane of odd width.
"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tools.cuemu import plugin  # noqa: E402

plugin.pytest_configure(None)
import numpy as np  # noqa: E402

from diligentfx_b200 import synth  # noqa: E402
from diligentfx_b200.chain import ChainConfig, PostProcessChain  # noqa: E402
from oracle import oracle_py as op  # noqa: E402


def run(tag, w, h, edit=lambda c: None, frames=3):
    from diligentfx_b200 import capi
    cfg = ChainConfig()
    edit(cfg)
    rev = bool(cfg.postfx_flags & capi.POSTFX_FLAG_REVERSED_DEPTH)
    try:
        seq = synth.generate_sequence(w, h, frames)
        if rev:
            seq = [synth.reverse_depth_frame(f) for f in seq]
        stages = op.STAGE_ALL
        if cfg.dof is not None:                                   # a lens that actually blurs the scene
            stages |= op.STAGE_DOF
            lens = []
            for f in seq:
                g = dict(f)
                for k in ("curr_camera", "prev_camera"):
                    c = capi.CameraAttribs.from_buffer_copy(bytes(f[k]))
                    c.fFocusDistance, c.fFStop = 6.0, 1.4
                    g[k] = c
                lens.append(g)
            seq = lens
        chain = PostProcessChain(w, h, cfg)
        o = op.Oracle(w, h)
        o.set_reversed_depth(rev)
        o.set_ssr(cfg.ssr, cfg.ssr_flags), o.set_ssao(cfg.ssao), o.set_ssao_flags(cfg.ssao_flags), o.set_bloom(cfg.bloom), o.set_taa(cfg.taa, cfg.taa_flags)
        o.set_tonemap(cfg.tonemap, cfg.ave_log_lum, cfg.to_srgb), o.set_compose_scales(cfg.ssr_scale, cfg.ssao_scale)
        if cfg.dof is not None:
            o.set_dof(cfg.dof, cfg.dof_flags)
        for fr in seq:
            ldr = chain.run_frame(fr).cpu().numpy()
            o.set_inputs(fr)
            o.frame(stages)
        d = np.abs(np.clip(ldr[..., :3], 0, 1).astype(np.float64) - np.clip(o.get("ldr")[..., :3], 0, 1))
        mse = (d ** 2).mean()
        print(f"{tag:44s} PSNR {200 if mse == 0 else 10 * np.log10(1 / mse):6.1f} dB  max {d.max():.1e}  finite={np.isfinite(ldr).all()}")
        chain.close()
    except Exception as e:  # a refusal (DfxError) is a result too
        print(f"{tag:44s} {type(e).__name__}: {str(e)[:150]}")
    finally:
        op.lib().orc_set_reversed_depth(0)


def flags(rev, ssr_flags, ssao_flags, dof, algo=0):
    from diligentfx_b200 import capi

    def f(c):
        c.postfx_flags = capi.POSTFX_FLAG_REVERSED_DEPTH if rev else 0
        c.ssr_flags, c.ssao_flags, c.ssao.Algorithm = ssr_flags, ssao_flags, algo
        if dof:
            a = capi.DOFAttribs.default()
            a.MaxCircleOfConfusion = 0.02
            c.dof, c.dof_flags = a, capi.DOF_FLAG_TEMPORAL_SMOOTHING | capi.DOF_FLAG_KARIS_INVERSE
    return f


def S(**kw):
    def f(c):
        for k, v in kw.items():
            obj, attr = k.split("__")
            setattr(getattr(c, obj), attr, v)
    return f


if __name__ == "__main__":
    for w, h in ((1, 1), (2, 2), (3, 5), (17, 9), (31, 33), (64, 1), (1, 64), (130, 70)):
        run(f"{w}x{h}", w, h)
    W, H = 97, 55
    for tag, edit in (("defaults", lambda c: None),
                      ("ssr MaxTraversalIntersections=0", S(ssr__MaxTraversalIntersections=0)), ("ssr MaxTraversalIntersections=1", S(ssr__MaxTraversalIntersections=1)),
                      ("ssr MaxTraversalIntersections=1000", S(ssr__MaxTraversalIntersections=1000)), ("ssr MostDetailedMip=2", S(ssr__MostDetailedMip=2)),
                      ("ssr MostDetailedMip=6", S(ssr__MostDetailedMip=6)), ("ssr RoughnessThreshold=0", S(ssr__RoughnessThreshold=0.0)),
                      ("ssr RoughnessThreshold=1", S(ssr__RoughnessThreshold=1.0)), ("ssr GGXImportanceSampleBias=1", S(ssr__GGXImportanceSampleBias=1.0)),
                      ("ssr SpatialReconstructionRadius=0", S(ssr__SpatialReconstructionRadius=0.0)), ("ssr BilateralSigma=0.3", S(ssr__BilateralCleanupSpatialSigmaFactor=0.3)),
                      ("ssao EffectRadius=0.01", S(ssao__EffectRadius=0.01)), ("ssao EffectRadius=50", S(ssao__EffectRadius=50.0)),
                      ("ssao SpatialReconstructionRadius=0", S(ssao__SpatialReconstructionRadius=0.0)), ("ssao TemporalStability=0", S(ssao__TemporalStabilityFactor=0.0)),
                      ("bloom Radius=0.2", S(bloom__Radius=0.2)), ("bloom Radius=1.0", S(bloom__Radius=1.0)), ("bloom Threshold=0", S(bloom__Threshold=0.0)),
                      ("taa TemporalStability=0", S(taa__TemporalStabilityFactor=0.0)), ("taa Reset every frame", S(taa__ResetAccumulation=1)),
                      ("tonemap white point 0.5", S(tonemap__fWhitePoint=0.5))):
        run(tag, W, H, edit)
    for rev, sf, af, dof in itertools.product((0, 1), (0, 1, 2, 3), (0, 1, 2, 3), (0, 1)):
        if not (dof and sf in (1, 3)):
            run(f"rev={rev} ssr_flags={sf} ssao_flags={af} dof={dof}", W, H, flags(rev, sf, af, dof))
    for algo, af in itertools.product((1, 2), (0, 2)):
        run(f"rev=1 ssao algorithm {algo} ssao_flags={af}", W, H, flags(1, 0, af, 0, algo))
