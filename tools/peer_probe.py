#!/usr/bin/env python3
"""Staged check of the NVLink peer-load path on 2+ GPUs (one process per GPU): CUDA-IPC mapping of the shared planes, a plain
peer read through the mapping, then the peer-sharded SSR against the gathered one. Every stage synchronises and reports, so
a failure names the stage.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_probe.py
"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from diligentfx_b200 import strips, synth  # noqa: E402
from diligentfx_b200.chain import INPUT_SPECS  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def say(*a):
        print(f"[rank {rank}]", *a, flush=True)

    def stage(name, fn):
        try:
            r = fn()
            torch.cuda.synchronize()
            say("ok  ", name, "" if r is None else r)
            return r
        except Exception:  # noqa: BLE001
            say("FAIL", name)
            traceback.print_exc()
            sys.stdout.flush()
            os._exit(3)

    W, H = 320, 256
    bounds = strips.strip_bounds(H, world)
    y0, y1 = bounds[rank]
    say("p2p matrix", [[torch.cuda.can_device_access_peer(i, j) for j in range(world) if j != i] for i in range(world)])
    mine = torch.full((H, W), float(rank + 1), device=dev)
    slab = stage("PeerSlab", lambda: strips.PeerSlab({"t": ((H, W), torch.float32)}, fill=float(rank + 1)))
    say("slab bases", [hex(b) for b in slab.base], "local mean", float(slab.local["t"].mean()))
    del mine

    seq = synth.generate_sequence(W, H, 2)
    r_peer = stage("runner(peer)", lambda: strips.SsrStripRunner(W, H, peer=True, poison=True))
    r_gath = stage("runner(gather)", lambda: strips.SsrStripRunner(W, H))
    for fr in seq:
        def mk():
            d = {}
            for n in INPUT_SPECS:
                full = torch.from_numpy(np.ascontiguousarray(fr[n])).to(dev)
                part = torch.full_like(full, float("nan"))
                part[y0:y1] = full[y0:y1]
                d[n] = part
            return d
        a = stage("gather execute", lambda: r_gath.execute(fr["frame"], mk(), fr["curr_camera"], fr["prev_camera"]).clone())
        b = stage("peer execute", lambda: r_peer.execute(fr["frame"], mk(), fr["curr_camera"], fr["prev_camera"]).clone())
        say("frame", fr["frame"], "identical:", bool(torch.equal(a[y0:y1], b[y0:y1])), "finite:", bool(torch.isfinite(b[y0:y1]).all()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
