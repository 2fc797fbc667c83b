#!/usr/bin/env python3
"""BASELINE.json config 4: ScreenSpaceReflection (S1-S7) on a 7680x4320 G-buffer, one frame split into 64-row-aligned strips
over the GPUs of the box, halo rows / gathered planes exchanged with NCCL send/recv over NVLink.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_ssr_strips.py [--steps K] [--warmup W]

Prints one JSON line (rank 0): Mpixels/s of SSR over the whole frame = W*H / max-over-ranks device time per frame, plus
the split between kernel time and exchange time of the slowest rank. Scaling is STRONG (the frame is fixed, strips shrink).
Synthetic inputs are generated per rank for its own rows only (the rest of every plane is filled by the exchanges).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from diligentfx_b200 import strips, synth  # noqa: E402
from diligentfx_b200.chain import INPUT_SPECS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=7680)
    ap.add_argument("--height", type=int, default=4320)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--balance", action="store_true", help="cost-balanced strip heights (reflective pixels weigh more) instead of equal ones")
    ap.add_argument("--peer", action="store_true", help="ray march with NVLink peer loads instead of gathering Hi-Z / colour / normal")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    W, H = a.width, a.height
    weights = None
    if a.balance and world > 1:
        # cost-balanced strips: fraction of reflection samples per row of a small rendering of the same frame
        small = synth.generate_frame(synth.make_scene(7), 0, 480, 270)
        refl = (small["material"][..., 0] <= 0.2) & (small["depth"] < 1.0)  # SSR defaults: roughness in .x, threshold 0.2
        weights = strips.reflective_block_cost(refl.mean(axis=1), H)
    bounds = strips.strip_bounds(H, world, weights=weights)
    y0, y1 = bounds[rank]
    # two consecutive frames of the camera path; only this rank's rows (+1 so that chunked generation lines up) are ray-cast
    scene = synth.make_scene(7)
    frames = []
    for f in (0, 1):
        cam, prev = synth.make_camera(f, W, H), synth.make_camera(max(f - 1, 0), W, H)
        full = {n: torch.zeros((H, W) + ((c,) if c else ()), dtype=torch.float32, device=dev) for n, c in INPUT_SPECS.items()}
        fr = synth.generate_rows(scene, f, W, H, y0, y1)
        for n in INPUT_SPECS:
            full[n][y0:y1] = torch.from_numpy(fr[n]).to(dev)
        frames.append((full, cam.attribs, prev.attribs))
    runner = strips.SsrStripRunner(W, H, peer=a.peer, input_sets=2, bounds=bounds)
    if runner.peer:  # the G-buffer lives in the runner's exported planes (double-buffered), as a renderer would write it
        for i, (full, _, _) in enumerate(frames):
            for n in ("depth", "color", "normal"):
                runner.shared_sets[i][n][y0:y1] = full[n][y0:y1]
                del full[n]

    def step(i):
        full, c, p = frames[i & 1]
        runner.execute(i, full, c, p, input_set=i & 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(a.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(a.warmup + i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = float(ms.item())
        print(json.dumps({"metric": "Mpixels/sec ScreenSpaceReflection Hi-Z ray-march @ 8K G-buffer, row-strip shard", "value": round(W * H / 1e6 / (t / 1e3), 2),
                          "unit": "Mpix/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(t, 4), "scaling": "strong",
                          "config": {"workload": f"SSR S1-S7 + PostFX prep, {W}x{H}, strips {bounds}", "exchange": ("NCCL send/recv of halo rows (4/1/24/4/1/24/2); ray march loads Hi-Z / colour / normal from the owning GPU over NVLink" if runner.peer
                                                  else "NCCL send/recv: halo rows (1/4/24/2) + gathered depth, colour, normal, Hi-Z")}}),
              flush=True)
    runner.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
