"""G-buffer frames on disk, for offline batches (BASELINE.json config 5): one `.npz` per frame holding the planes in the
formats the renderer stores them in (Hydrogent/src/Tasks/HnBeginFrameTask.cpp:63-69 - colour / normal RGBA16F, motion RG16F, material RG8,
depth D32_FLOAT = 26 B/px before compression) plus both cameras as the raw bytes of `CameraAttribs` (BasicStructures.fxh layout) and the
frame index. What `load_frame` returns is what `PostProcessChain.stream_frames(packed=True)` consumes; `save_ldr` / `load_ldr` carry the
RGBA8 result (`.npy`, or binary PPM for a quick look). No image library is involved: NumPy's container only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import capi
from .chain import PACKED_SPECS, pack_frame

FORMAT_VERSION = 1
_PLANES = {"color16": (np.float16, 4), "normal16": (np.float16, 4), "motion16": (np.float16, 2), "material8": (np.uint8, 2), "depth": (np.float32, 0)}


def _camera_bytes(cam) -> np.ndarray:
    return np.frombuffer(bytes(cam), np.uint8).copy()


def _camera_from(buf: np.ndarray):
    if buf.size != C.sizeof(capi.CameraAttribs):
        raise ValueError(f"camera block of {buf.size} bytes, expected {C.sizeof(capi.CameraAttribs)}")
    return capi.CameraAttribs.from_buffer_copy(buf.tobytes())


def save_frame(path: str, frame: dict, compressed: bool = False) -> None:
    """`frame`: a frame dict as `synth.generate_sequence` makes them (fp32 planes) or one already packed by `chain.pack_frame`.
    The previous depth is not stored: it is the depth of the frame before (HnPostProcessTask.cpp:788-832)."""
    p = frame if all(s[0] in frame for s in PACKED_SPECS.values()) else pack_frame(frame)
    arrays = {}
    for key, (dt, ch) in _PLANES.items():
        a = p[key].numpy() if isinstance(p[key], torch.Tensor) else np.asarray(p[key])
        if a.dtype != dt or (ch and a.shape[-1] != ch):
            raise ValueError(f"plane {key}: {a.dtype} {a.shape}, expected {np.dtype(dt)} with {ch or 1} channel(s)")
        arrays[key] = np.ascontiguousarray(a)
    h, w = arrays["depth"].shape
    for key, a in arrays.items():
        if a.shape[:2] != (h, w):
            raise ValueError(f"plane {key} is {a.shape[1]}x{a.shape[0]}, depth is {w}x{h}")
    meta = np.array([FORMAT_VERSION, int(frame["frame"]), w, h], np.int64)
    (np.savez_compressed if compressed else np.savez)(path, meta=meta, curr_camera=_camera_bytes(frame["curr_camera"]), prev_camera=_camera_bytes(frame["prev_camera"]), **arrays)


def load_frame(path: str, pin: bool = False) -> dict:
    """The packed frame dict of `path` (tensors share no memory with the file; `pin=True` puts them in page-locked memory for the
    streaming pipeline). No "prev_depth" entry: `stream_frames` keeps the previous frame's depth on the device."""
    with np.load(path) as z:
        meta = z["meta"]
        if int(meta[0]) != FORMAT_VERSION:
            raise ValueError(f"{path}: format version {int(meta[0])}, this build reads {FORMAT_VERSION}")
        out = {"frame": int(meta[1]), "curr_camera": _camera_from(z["curr_camera"]), "prev_camera": _camera_from(z["prev_camera"])}
        w, h = int(meta[2]), int(meta[3])
        for key, (dt, ch) in _PLANES.items():
            a = z[key]
            if a.dtype != dt or a.shape != ((h, w, ch) if ch else (h, w)):
                raise ValueError(f"{path}: plane {key} is {a.dtype} {a.shape}")
            t = torch.from_numpy(np.ascontiguousarray(a))
            out[key] = t.pin_memory() if pin else t
    return out


def sequence_paths(directory: str) -> list[str]:
    """The `.npz` frames of a directory in frame order (by the stored index, then by name)."""
    names = sorted(n for n in os.listdir(directory) if n.endswith(".npz"))
    keyed = []
    for n in names:
        with np.load(os.path.join(directory, n)) as z:
            keyed.append((int(z["meta"][1]), n))
    return [os.path.join(directory, n) for _, n in sorted(keyed)]


def save_ldr(path: str, ldr) -> None:
    """RGBA8 result of the chain ((H, W, 4) uint8, or float in [0, 1]) as `.npy`, or as binary PPM (RGB) when `path` ends in `.ppm`."""
    a = ldr.cpu().numpy() if isinstance(ldr, torch.Tensor) else np.asarray(ldr)
    if a.dtype != np.uint8:
        a = np.floor(np.clip(a.astype(np.float32), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    if path.endswith(".ppm"):
        with open(path, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (a.shape[1], a.shape[0]))
            f.write(np.ascontiguousarray(a[..., :3]).tobytes())
    else:
        np.save(path, a)


def load_ldr(path: str) -> np.ndarray:
    if path.endswith(".ppm"):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"P6"
            w, h = (int(v) for v in f.readline().split())
            assert int(f.readline()) == 255
            return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)
    return np.load(path)
