#!/usr/bin/env python3
"""Regenerates tests/golden/reference_shaders_49x27.npz: the outputs of the REFERENCE'S OWN pixel shaders (compiled for the
CPU by oracle/refshader from /root/reference/Shaders, which must be mounted) for every pass of the third frame of a
synthetic 49 x 27 sequence, each pass fed with the oracle's planes as inputs - default chain, then DepthOfField.

Unlike chain_96x54.npz (outputs of the oracle), these vectors come from the reference: tests/test_reference_shader_golden.py
holds the oracle to them bit for bit on machines where neither /root/reference nor librefshaders.so exists.

    python tests/golden/make_reference_shader_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from diligentfx_b200 import capi, synth  # noqa: E402

W, H, WARM = 49, 27, 2
OUT = os.path.join(HERE, "reference_shaders_49x27.npz")


def configs():
    """(prefix, variant, frames): the benchmarked chain, and the chain with DepthOfField between TAA and Bloom."""
    from oracle.refshader.driver import Variant
    seq = synth.generate_sequence(W, H, WARM + 1)
    a = capi.DOFAttribs.default()
    a.MaxCircleOfConfusion = 0.02
    lens = []
    for f in seq:
        g = dict(f)
        for k in ("curr_camera", "prev_camera"):
            c = capi.CameraAttribs.from_buffer_copy(bytes(f[k]))
            c.fFocusDistance, c.fFStop = 6.0, 1.4
            g[k] = c
        lens.append(g)
    return [("chain", Variant(), seq),
            ("dof", Variant(dof=True, dof_flags=capi.DOF_FLAG_TEMPORAL_SMOOTHING | capi.DOF_FLAG_KARIS_INVERSE, dof_attribs=a), lens)]


def run(with_reference: bool) -> dict:
    """{"<prefix>/<pass label>": plane}: reference-shader outputs (with_reference) or the oracle's outputs for the same passes."""
    from oracle.refshader.driver import compare_frame, make_oracle
    out = {}
    for prefix, v, frames in configs():
        o = make_oracle(W, H, v)
        for fr in frames[:WARM]:
            o.set_inputs(fr)
            o.frame(v.stages())
        for label, (got, want) in compare_frame(o, frames[WARM], v, with_reference).items():
            if prefix == "dof" and label[0] != "D":
                continue                                              # the rest repeats the first configuration
            out[f"{prefix}/{label}"] = got if with_reference else want
    return out


if __name__ == "__main__":
    from oracle.refshader import refsh
    refsh.build()
    assert refsh.available(), "needs /root/reference (or a prebuilt oracle/_ref/librefshaders.so)"
    planes = run(True)
    np.savez_compressed(OUT, **planes)
    print(f"{OUT}: {len(planes)} planes, {os.path.getsize(OUT) / 1024:.0f} KiB")
