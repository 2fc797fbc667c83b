#!/usr/bin/env python3
"""Run a few frames of the chain at a given size (profiling target for ncu: few launches, no timing)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diligentfx_b200 import synth  # noqa: E402
from diligentfx_b200.chain import PostProcessChain  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--frames", type=int, default=3)
a = ap.parse_args()
seq = synth.generate_sequence(a.width, a.height, a.frames)
chain = PostProcessChain(a.width, a.height)
for fr in seq:
    chain.run_frame(fr)
torch.cuda.synchronize()
print("done", chain.lib.dfx_launch_count(), "launches")
