"""cuemu — DEVELOPMENT TOOL (not collected by a plain `pytest tests/`). The 2-rank row-strip SSR of tests/test_strips_gpu.py on the HOST
build of the kernels: two processes, gloo instead of NCCL, CPU tensors instead of device planes; for the peer-load variant the
slabs are shared-memory segments mapped into the other process (CUEMU_IPC), standing in for CUDA IPC over NVLink. Each rank's strip must be
bit-identical to the single-process run of the same kernels.

    python tests/cuemu_strips_check.py
"""
import os
import socket
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank: int, world: int, port: int, w: int, h: int, frames: int, out_dir: str, peer: bool):
    os.environ.setdefault("CUEMU_THREADS", "16")
    if peer:
        os.environ["CUEMU_IPC"] = "1"                                # device allocations become shared-memory segments the other rank can map
    from tools.cuemu import plugin
    plugin.pytest_configure(None)
    import torch
    import torch.distributed as dist

    from diligentfx_b200 import synth
    from diligentfx_b200.chain import INPUT_SPECS, STAGE_POSTFX, STAGE_SSR, ChainConfig, PostProcessChain
    from diligentfx_b200.strips import SsrStripRunner, strip_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seq = synth.generate_sequence(w, h, frames)
        bounds = strip_bounds(h, world, weights=[1.0, 3.0, 3.0, 3.0][:h // 64] if peer else None)   # peer run: unequal, cost-balanced strips
        y0, y1 = bounds[rank]
        runner = SsrStripRunner(w, h, peer=peer, poison=True, bounds=bounds, device=torch.device("cpu"))
        ref = PostProcessChain(w, h, ChainConfig(stages=STAGE_POSTFX | STAGE_SSR)) if rank == 0 else None
        for fr in seq:
            inputs = {}
            for n in INPUT_SPECS:
                full = torch.from_numpy(np.ascontiguousarray(fr[n])).clone()
                part = torch.full_like(full, float("nan"))           # rows this rank does not own are poison until exchanged
                part[y0:y1] = full[y0:y1]
                inputs[n] = part
            out = runner.execute(fr["frame"], inputs, fr["curr_camera"], fr["prev_camera"])
            if ref is not None:
                ref.run_frame(fr)
        np.save(os.path.join(out_dir, f"strip_{rank}.npy"), out[y0:y1].cpu().numpy())
        runner.close()
        if ref is not None:
            np.save(os.path.join(out_dir, "ref.npy"), ref.fetch("ssr", 0))
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp

    from diligentfx_b200.strips import strip_bounds
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    w, h, world = 320, 256, 2
    for peer in (False, True):
        with tempfile.TemporaryDirectory() as tmp:
            mp.spawn(_worker, args=(world, port + int(peer), w, h, 3, tmp, peer), nprocs=world, join=True)
            ref = np.load(os.path.join(tmp, "ref.npy"))
            for r, (y0, y1) in enumerate(strip_bounds(h, world, weights=[1.0, 3.0, 3.0, 3.0] if peer else None)):
                got = np.load(os.path.join(tmp, f"strip_{r}.npy"))
                assert np.isfinite(got).all(), "poison rows leaked into the owned strip: an exchange is missing"
                assert np.array_equal(got, ref[y0:y1]), f"strip {r} differs from the single-process result (max abs {np.abs(got - ref[y0:y1]).max()})"
                print(f"{'peer loads' if peer else 'gathered planes'}: rank {r}, rows [{y0}, {y1}) bit-identical to the single-process run")
