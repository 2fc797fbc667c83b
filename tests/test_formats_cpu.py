"""Host side of the transfer-format path (diligentfx_b200.chain.pack_frame / widen_frame / pack_ldr8): no GPU needed."""
import numpy as np

from diligentfx_b200 import synth
from diligentfx_b200.chain import PACKED_SPECS, pack_frame, pack_ldr8, widen_frame


def test_pack_widen_round_trip():
    fr = synth.generate_sequence(64, 36, 1)[0]
    p = pack_frame(fr)
    assert p["color16"].numpy().dtype == np.float16 and p["color16"].shape == (36, 64, 4)
    assert p["motion16"].shape == (36, 64, 2) and p["material8"].numpy().dtype == np.uint8 and p["material8"].shape == (36, 64, 2)
    assert not any(k in p for k in PACKED_SPECS)  # only the narrow planes travel
    w = widen_frame(p)
    for name in ("color", "normal", "motion"):
        ref = np.asarray(fr[name], np.float32)
        assert w[name].dtype == np.float32 and w[name].shape == ref.shape
        assert np.all(np.abs(w[name] - ref) <= np.abs(ref) * 2.0 ** -11 + 6e-8), name  # half rounding: 11 significant bits
    assert np.all(np.abs(w["material"][..., :2] - np.clip(fr["material"][..., :2], 0, 1)) <= 0.5 / 255 + 1e-7)
    assert np.all(w["material"][..., 2:] == 0)
    assert np.array_equal(w["depth"], fr["depth"]) and np.array_equal(w["prev_depth"], fr["prev_depth"])
    # widening is idempotent: packing the widened frame again changes nothing
    p2 = pack_frame(w)
    for key, _, _ in PACKED_SPECS.values():
        assert np.array_equal(p2[key].numpy(), p[key].numpy()), key
    # bytes per pixel that cross PCIe: 8 + 8 + 4 + 2 (+ 4 + 4 for the two depths)
    assert sum(p[k].element_size() * p[k].shape[-1] for k, _, _ in PACKED_SPECS.values()) == 22


def test_pack_ldr8_rule():
    v = np.array([np.nan, -1.0, 0.0, 0.5 / 255, 1.5 / 255, 0.999, 1.0, 7.0, 127.5 / 255], np.float32)
    assert pack_ldr8(v).tolist() == [0, 0, 0, 1, 2, 255, 255, 255, 128]
