#!/usr/bin/env python3
"""Regenerates tests/golden/chain_96x54.npz — outputs of the ORACLE for 3 consecutive synthetic 96x54 frames.

The reference has no golden vectors for this path and cannot run here, so these fixtures pin the
*oracle*, not the reference: they make any change to the oracle's arithmetic visible in review (the CPU test compares
bit-exactly) and give the GPU tests a target that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diligentfx_b200 import synth  # noqa: E402
from oracle import oracle_py as op  # noqa: E402

W, H, FRAMES = 96, 54, 3
PLANES = ["reproj_depth", "closest_motion", "ssao_occ", "ssao_out", "ssr_out", "composed", "bloom_out", "ldr"]


def run():
    seq = synth.generate_sequence(W, H, FRAMES)
    o = op.Oracle(W, H, threads=1)
    for fr in seq:
        o.set_inputs(fr)
        o.frame()
    out = {p: o.get(p) for p in PLANES}
    out["taa_accum"] = o.get(f"taa_accum{(FRAMES - 1) & 1}")
    out["ssao_histlen"] = o.get(f"ssao_histlen{(FRAMES - 1) & 1}")
    return out


def run_variants():
    """The feature-flag variants and the passes outside the benchmarked chain, two frames each."""
    from diligentfx_b200 import capi
    seq = synth.generate_sequence(W, H, 2)
    out = {}

    def frames(o, frs, stages=op.STAGE_ALL):
        for fr in frs:
            o.set_inputs(fr)
            o.frame(stages)

    o = op.Oracle(W, H, threads=1)                                       # reversed depth
    o.set_reversed_depth(True)
    try:
        frames(o, [synth.reverse_depth_frame(f) for f in seq])
        out["reversed_ldr"], out["reversed_ssr"] = o.get("ldr"), o.get("ssr_out")
    finally:
        o.set_reversed_depth(False)
    o = op.Oracle(W, H, threads=1)                                       # half-resolution SSAO and SSR
    o.set_ssao_flags(capi.SSAO_FLAG_HALF_RESOLUTION)
    o.set_ssr(capi.SSRAttribs.default(), capi.SSR_FLAG_HALF_RESOLUTION)
    frames(o, seq)
    out["half_ssao"], out["half_ssr"], out["half_ldr"] = o.get("ssao_out"), o.get("ssr_out"), o.get("ldr")
    o = op.Oracle(W, H, threads=1)                                       # DepthOfField between TAA and Bloom
    a = capi.DOFAttribs.default()
    a.MaxCircleOfConfusion = 0.02
    o.set_dof(a, capi.DOF_FLAG_TEMPORAL_SMOOTHING | capi.DOF_FLAG_KARIS_INVERSE)
    lens = []
    for f in seq:
        g = dict(f)
        for k in ("curr_camera", "prev_camera"):
            c = capi.CameraAttribs.from_buffer_copy(bytes(f[k]))
            c.fFocusDistance, c.fFStop = 6.0, 1.4
            g[k] = c
        lens.append(g)
    frames(o, lens, op.STAGE_ALL | op.STAGE_DOF)
    out["dof_out"], out["dof_ldr"] = o.get("dof_out"), o.get("ldr")
    o.brdf_lut(32, 128)                                                  # pre-integrated GGX table + full compose on the same frame
    rng = np.random.default_rng(9)
    o.set("base_color", np.concatenate([rng.uniform(0.02, 1.0, (H, W, 3)), np.ones((H, W, 1))], -1).astype(np.float32))
    o.set("specular_ibl", np.concatenate([np.exp2(rng.uniform(-4, 2, (H, W, 3))), np.ones((H, W, 1))], -1).astype(np.float32))
    o.set_compose_scales(0.8, 0.6)
    o.run("compose_ibl")
    out["brdf_lut"], out["composed_ibl"] = o.get("brdf_lut"), o.get("composed")
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, fn in (("chain_96x54.npz", run), ("variants_96x54.npz", run_variants)):
        out = fn()
        path = os.path.join(here, name)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes", {k: v.shape for k, v in out.items()})
