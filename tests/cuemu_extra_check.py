"""cuemu — DEVELOPMENT TOOL. Kernel-source checks that the standard parity tests do not reach, run on the host build:

    python -m pytest -p tools.cuemu.plugin tests/cuemu_extra_check.py -q   (not collected by a plain `pytest tests/`: the file name does not match test_*.py)
"""
import ctypes as C

import numpy as np
import pytest

from diligentfx_b200 import capi, synth
from helpers import Dev, rows


@pytest.mark.parametrize("size", [(34, 18), (33, 17)], ids=["even", "odd"])
def test_bilateral_filter_on_every_pixel_including_odd_edges(size):
    """ssr_bilateral_kernel with the filter branch live everywhere (cf. tests/test_reference_shaders.py::
    test_bilateral_quad_derivatives_everywhere, which holds the oracle to the reference shader on the same planes)."""
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
    rough = o.get("ssr_roughness")
    ci = fr["frame"] & 1
    rad, var = rng.uniform(0, 4, (h, w, 4)).astype(np.float32), np.ones((h, w), np.float32)
    o.set(f"ssr_radhist{ci}", rad), o.set(f"ssr_varhist{ci}", var)
    o.run("ssr_bilateral")
    want = o.get("ssr_out")
    d = Dev()
    out = d.empty(h, w, 4, fill=5.0)
    P = lambda a: C.byref(d.plane(d.up(a)))  # noqa: E731
    capi.check(d.lib.dfx_pass_ssr_bilateral(None, d.cameras(fr["curr_camera"], fr["prev_camera"]), C.byref(ssr), C.byref(d.plane(d.mask(np.ones((h, w))))),
                                            P(depth), P(normal), P(rough), P(rad), P(var), C.byref(d.plane(out)), rows(h)))
    got = d.host(out)
    rel = np.abs(got - want) / (1.0 + np.abs(want))
    edge = np.zeros((h, w), bool)
    edge[:, -1] = edge[-1, :] = True
    assert rel.max() < 1e-4, (rel.max(), rel[edge].max())
