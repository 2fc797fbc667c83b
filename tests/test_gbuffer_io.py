"""G-buffer frame files (diligentfx_b200/gbuffer_io.py): what is written is what the streaming pipeline would be handed - CPU only."""
import ctypes as C
import os

import numpy as np

from diligentfx_b200 import capi, gbuffer_io, synth
from diligentfx_b200.chain import PACKED_SPECS, pack_frame


def test_frame_round_trip(tmp_path):
    seq = synth.generate_sequence(96, 54, 3)
    for compressed in (False, True):
        d = tmp_path / ("z" if compressed else "raw")
        os.makedirs(d)
        for fr in reversed(seq):   # written out of order: sequence_paths sorts by the stored frame index
            gbuffer_io.save_frame(str(d / f"f{fr['frame']:04d}.npz"), fr, compressed=compressed)
        paths = gbuffer_io.sequence_paths(str(d))
        assert len(paths) == 3
        for fr, path in zip(seq, paths):
            got, want = gbuffer_io.load_frame(path), pack_frame(fr)
            assert got["frame"] == fr["frame"] and "prev_depth" not in got
            for key, _, _ in PACKED_SPECS.values():
                assert np.array_equal(got[key].numpy(), want[key].numpy()), key
            assert np.array_equal(got["depth"].numpy(), np.asarray(fr["depth"], np.float32))
            for cam in ("curr_camera", "prev_camera"):
                assert bytes(got[cam]) == bytes(fr[cam]) and C.sizeof(got[cam]) == C.sizeof(capi.CameraAttribs)
    raw = sum(os.path.getsize(p) for p in gbuffer_io.sequence_paths(str(tmp_path / "raw")))
    assert raw < 3 * (96 * 54 * 26 + 8192)   # 26 B/px in the renderer's formats + cameras + container


def test_packed_frames_are_accepted_and_bad_files_rejected(tmp_path):
    fr = synth.generate_sequence(64, 36, 1)[0]
    p = str(tmp_path / "a.npz")
    gbuffer_io.save_frame(p, pack_frame(fr) | {"frame": 7})
    assert gbuffer_io.load_frame(p)["frame"] == 7
    z = dict(np.load(p))
    z["meta"] = np.array([99, 0, 64, 36], np.int64)
    np.savez(str(tmp_path / "b.npz"), **z)
    try:
        gbuffer_io.load_frame(str(tmp_path / "b.npz"))
    except ValueError as e:
        assert "version" in str(e)
    else:
        raise AssertionError("a file of another format version must be refused")


def test_ldr_files(tmp_path):
    rng = np.random.default_rng(5)
    f = rng.random((20, 30, 4), dtype=np.float32)
    gbuffer_io.save_ldr(str(tmp_path / "o.npy"), f)
    a = gbuffer_io.load_ldr(str(tmp_path / "o.npy"))
    assert a.dtype == np.uint8 and a.shape == (20, 30, 4) and np.abs(a.astype(np.float32) / 255.0 - f).max() <= 0.5 / 255.0 + 1e-6
    gbuffer_io.save_ldr(str(tmp_path / "o.ppm"), a)
    assert np.array_equal(gbuffer_io.load_ldr(str(tmp_path / "o.ppm")), a[..., :3])
