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


if __name__ == "__main__":
    out = run()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chain_96x54.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", {k: v.shape for k, v in out.items()})
