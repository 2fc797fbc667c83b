"""Shared helpers for the parity tests: device plane plumbing over the pass-level C-ABI and error metrics."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from diligentfx_b200 import capi  # noqa: E402


def psnr(a: np.ndarray, b: np.ndarray, peak: float = 1.0) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = float(np.mean((a - b) ** 2))
    return 200.0 if mse == 0.0 else 10.0 * np.log10(peak * peak / mse)


def reinhard(x: np.ndarray) -> np.ndarray:
    x = np.maximum(np.asarray(x, np.float64), 0.0)
    return x / (1.0 + x)


def outlier_fraction(a: np.ndarray, b: np.ndarray, tol: float) -> float:
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    if d.ndim == 3:
        d = d.max(axis=2)
    return float((d > tol).mean())


def assert_close(name: str, got: np.ndarray, want: np.ndarray, *, tol: float, max_outliers: float = 0.0, min_psnr: float | None = None,
                 hdr: bool = False, mask: np.ndarray | None = None):
    """|got-want| <= tol everywhere except at most `max_outliers` fraction of pixels (branch flips at discontinuities);
    optional PSNR floor (peak 1; HDR planes are compared after Reinhard c/(1+c))."""
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    g, w = (reinhard(got), reinhard(want)) if hdr else (np.asarray(got, np.float64), np.asarray(want, np.float64))
    g = np.nan_to_num(g, nan=0.0, posinf=1e30, neginf=-1e30)
    w = np.nan_to_num(w, nan=0.0, posinf=1e30, neginf=-1e30)
    if mask is not None:
        m = mask if g.ndim == 2 else mask[..., None]
        g, w = g * m, w * m
    frac = outlier_fraction(g, w, tol)
    p = psnr(g, w)
    msg = f"{name}: outliers(>{tol:g})={frac:.5%} (allowed {max_outliers:.5%}), max_abs={np.abs(g - w).max():.3e}, psnr={p:.1f} dB"
    assert frac <= max_outliers, msg
    if min_psnr is not None:
        assert p >= min_psnr, msg
    return msg


# ---------------------------------------------------------------------------------------------------------------------
# device plumbing
# ---------------------------------------------------------------------------------------------------------------------
class Dev:
    """Keeps torch tensors alive while their dfx_plane descriptors are in flight."""

    def __init__(self):
        import torch
        self.torch = torch
        self.lib = capi.load()
        self.keep = []

    def up(self, a: np.ndarray, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(a, np.float32 if dtype is None else dtype)).cuda()
        self.keep.append(t)
        return t

    def mask(self, a: np.ndarray):
        t = self.torch.from_numpy(np.ascontiguousarray(a != 0, np.uint8)).cuda()
        self.keep.append(t)
        return t

    def empty(self, h: int, w: int, ch: int = 1, fill: float | None = None, dtype=None):
        shape = (h, w) if ch == 1 else (h, w, ch)
        dt = dtype or self.torch.float32
        t = self.torch.empty(shape, dtype=dt, device="cuda") if fill is None else self.torch.full(shape, fill, dtype=dt, device="cuda")
        self.keep.append(t)
        return t

    def plane(self, t) -> capi.Plane:
        p = capi.plane_of(t)
        self.keep.append(p)
        return p

    def pyr(self, tensors) -> capi.Pyramid:
        p = capi.pyramid_of(tensors)
        self.keep.append(p)
        return p

    def cameras(self, curr, prev):
        buf = (capi.CameraAttribs * 2)(curr, prev)
        t = self.torch.frombuffer(bytearray(bytes(buf)), dtype=self.torch.uint8).cuda()
        self.keep.append(t)
        return C.c_void_p(t.data_ptr())

    def sync(self):
        self.torch.cuda.synchronize()

    @staticmethod
    def host(t) -> np.ndarray:
        return t.detach().cpu().numpy().astype(np.float32)


def rows(h: int) -> capi.Rows:
    return capi.Rows(0, h)
