"""Row-strip sharded SSR on 2 GPUs (NCCL send/recv of halo rows / gathered planes) must be BIT-IDENTICAL to the single-GPU run
of the same kernels (SURVEY.md §8e parity requirement). Needs >= 2 CUDA devices: run with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, w: int, h: int, frames: int, out_dir: str, peer: bool):
    import torch
    import torch.distributed as dist

    from diligentfx_b200 import synth
    from diligentfx_b200.chain import INPUT_SPECS, STAGE_POSTFX, STAGE_SSR, ChainConfig, PostProcessChain
    from diligentfx_b200.strips import SsrStripRunner, strip_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        seq = synth.generate_sequence(w, h, frames)
        bounds = strip_bounds(h, world, weights=[1.0, 3.0, 3.0, 3.0][:h // 64] if peer else None)  # peer run: unequal, cost-balanced strips
        y0, y1 = bounds[rank]
        runner = SsrStripRunner(w, h, peer=peer, poison=True, bounds=bounds)
        ref = PostProcessChain(w, h, ChainConfig(stages=STAGE_POSTFX | STAGE_SSR)) if rank == 0 else None
        for fr in seq:
            inputs = {}
            for n in INPUT_SPECS:
                full = torch.from_numpy(np.ascontiguousarray(fr[n])).cuda()
                part = torch.full_like(full, float("nan"))  # rows this rank does not own are poison until exchanged
                part[y0:y1] = full[y0:y1]
                inputs[n] = part
            out = runner.execute(fr["frame"], inputs, fr["curr_camera"], fr["prev_camera"])
            if ref is not None:
                ref.run_frame(fr)
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"strip_{rank}.npy"), out[y0:y1].cpu().numpy())
        runner.close()
        if ref is not None:
            np.save(os.path.join(out_dir, "ref.npy"), ref.fetch("ssr", 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer", [False, True], ids=["nccl-gather", "nvlink-peer-loads"])
def test_ssr_strips_bit_identical_on_two_gpus(built, tmp_path, peer):
    import torch
    import torch.multiprocessing as mp

    from diligentfx_b200.strips import strip_bounds
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    if peer and not torch.cuda.can_device_access_peer(0, 1):
        pytest.skip("GPUs 0 and 1 have no peer path")
    w, h, world = 320, 256, 2
    mp.spawn(_worker, args=(world, _free_port(), w, h, 3, str(tmp_path), peer), nprocs=world, join=True)
    ref = np.load(tmp_path / "ref.npy")
    for r, (y0, y1) in enumerate(strip_bounds(h, world, weights=[1.0, 3.0, 3.0, 3.0] if peer else None)):
        got = np.load(tmp_path / f"strip_{r}.npy")
        assert np.isfinite(got).all(), "poison rows leaked into the owned strip: an exchange is missing"
        assert np.array_equal(got, ref[y0:y1]), f"strip {r} differs from the single-GPU result (max abs {np.abs(got - ref[y0:y1]).max()})"
