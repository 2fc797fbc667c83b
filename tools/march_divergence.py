"""Divergence of the Hi-Z march over neighbouring rays, from the oracle's per-pixel trip counts (test infrastructure used for analysis only):
mean trips per ray vs the mean over SIMD groups of the group maximum, for several 32-pixel group shapes. The ratio bounds what re-packing
rays inside a warp could gain.   usage: python tools/march_divergence.py [W H [frames]]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diligentfx_b200 import synth
from oracle import oracle_py as op

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2
seq = synth.generate_sequence(W, H, frames)
o = op.Oracle(W, H, threads=os.cpu_count() or 1)
plane = np.zeros((H, W), np.uint16)
for fr in seq:
    plane[:] = 0
    op.lib().orc_march_iteration_plane(plane.ctypes.data_as(C.POINTER(C.c_uint16)), W)
    o.set_inputs(fr)
    o.frame()
op.lib().orc_march_iteration_plane(None, 0)
it = plane.astype(np.float64)
traced = it > 0
print(f"{W}x{H}: rays {traced.mean():.3f} of the pixels, trips per ray mean {it[traced].mean():.1f} median {np.median(it[traced]):.0f} p90 {np.percentile(it[traced], 90):.0f} max {it.max():.0f}")
print(f"trips per PIXEL (untraced = 0): {it.mean():.2f}")
for gw, gh in ((32, 1), (16, 2), (8, 4), (4, 8)):
    h2, w2 = H // gh * gh, W // gw * gw
    g = it[:h2, :w2].reshape(h2 // gh, gh, w2 // gw, gw).transpose(0, 2, 1, 3).reshape(-1, gw * gh)
    gmax = g.max(axis=1)
    print(f"  group {gw:2d}x{gh}: mean of group max {gmax.mean():.2f} trips per pixel slot -> lane utilisation of the loop {g.mean() / gmax.mean():.3f}; groups with no ray {np.mean(gmax == 0):.3f}")
