#!/usr/bin/env python3
"""Runs the 4K chain eagerly for a few steady-state frames, then jumps the frame index (every history resets) and runs three more frames
between cudaProfilerStart / Stop: under `ncu --profile-from-start off` the capture holds one steady-state frame followed by the frames in
which the temporal filters are live (SSAO resampling / spatial reconstruction run their taps). Usage: see tools/gpu_r2c.sh."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diligentfx_b200 import synth  # noqa: E402
from diligentfx_b200.chain import INPUT_SPECS, PACKED_SPECS, ChainConfig, PostProcessChain, pack_frame, widen_frame  # noqa: E402

W, H = 3840, 2160
seq = synth.generate_sequence(W, H, 2)
packed = [pack_frame(fr) for fr in seq]
wide = [widen_frame(p) for p in packed]
res = [{n: (packed[i][PACKED_SPECS[n][0]].cuda() if n in PACKED_SPECS else torch.from_numpy(np.ascontiguousarray(wide[i][n])).cuda()) for n in INPUT_SPECS} for i in range(2)]
chain = PostProcessChain(W, H, ChainConfig(graph=False, overlap=False))
idx = 0
for _ in range(8):
    chain.execute(idx, seq[idx & 1]["curr_camera"], seq[idx & 1]["prev_camera"], res[idx & 1])
    idx += 1
torch.cuda.synchronize()
torch.cuda.profiler.start()
chain.execute(idx, seq[idx & 1]["curr_camera"], seq[idx & 1]["prev_camera"], res[idx & 1])      # steady state
idx += 100                                                                                        # history reset
for _ in range(3):
    chain.execute(idx, seq[idx & 1]["curr_camera"], seq[idx & 1]["prev_camera"], res[idx & 1])  # filters live
    idx += 1
torch.cuda.synchronize()
torch.cuda.profiler.stop()
chain.close()
