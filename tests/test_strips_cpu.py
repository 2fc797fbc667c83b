"""Row-strip partitioning and exchange logic on CPU tensors over gloo (world_size 2 and 3): the N > 1 host path of
diligentfx_b200/strips.py without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diligentfx_b200.strips import exchange_halo, gather_rows, strip_bounds


def test_strip_bounds_are_aligned_and_cover():
    for h, n in ((4320, 8), (4320, 4), (4320, 2), (2160, 8), (1080, 8), (187, 2), (64, 3)):
        b = strip_bounds(h, n)
        assert b[0][0] == 0 and b[-1][1] == h and len(b) == n
        for (a0, a1), (b0, b1) in zip(b, b[1:]):
            assert a1 == b0
        for y0, y1 in b:
            assert y0 % 64 == 0 and (y1 % 64 == 0 or y1 == h) and y1 >= y0
    assert strip_bounds(4320, 8)[0] == (0, 576)  # ceil(67.5 blocks / 8): the first strips take 9 blocks of 64 rows


def test_cost_balanced_strip_bounds():
    from diligentfx_b200.strips import reflective_block_cost
    # a frame whose lower 60 % is reflective: equal-height strips would give the last rank ~12x the work of the first
    h = 4320
    frac = [0.0] * 108 + [1.0] * 162  # per row of a 270-row preview
    w = reflective_block_cost(frac, h)
    assert len(w) == 68 and w[0] == 64.0 and abs(w[40] - 64 * 12.0) < 1e-9
    for n in (2, 3, 4, 8):
        b = strip_bounds(h, n, weights=w)
        assert b[0][0] == 0 and b[-1][1] == h and len(b) == n
        assert all(a1 == b0 for (_, a1), (b0, _) in zip(b, b[1:]))
        assert all(y0 % 64 == 0 and y1 > y0 for y0, y1 in b)
        cost = [sum(w[y0 // 64:-(-y1 // 64)]) for y0, y1 in b]
        assert max(cost) <= 1.25 * sum(cost) / n, (n, b, cost)  # within a block's worth of the ideal share
        equal = strip_bounds(h, n)
        assert max(cost) < max(sum(w[y0 // 64:-(-y1 // 64)]) for y0, y1 in equal)
    assert strip_bounds(128, 2, weights=[0.0, 5.0]) == [(0, 64), (64, 128)]  # nobody is left without a block
    assert strip_bounds(64, 3, weights=[1.0])[-1][1] == 64  # fewer blocks than ranks: empty strips are allowed


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, h: int, w: int):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bounds = strip_bounds(h, world)
        y0, y1 = bounds[rank]
        truth = torch.arange(h * w * 2, dtype=torch.float32).reshape(h, w, 2)
        truth1 = torch.arange(h * w, dtype=torch.float32).reshape(h, w) * 0.5

        def mine(t):
            p = torch.full_like(t, -1.0)
            p[y0:y1] = t[y0:y1]
            return p

        # halo exchange: afterwards own rows +- halo match the truth, everything else is untouched
        for halo in (1, 4, 24):
            a, b = mine(truth), mine(truth1)
            exchange_halo([a, b], bounds, halo)
            lo, hi = max(y0 - halo, 0), min(y1 + halo, h)
            if y1 > y0:
                assert torch.equal(a[lo:hi], truth[lo:hi]) and torch.equal(b[lo:hi], truth1[lo:hi]), (rank, halo)
                assert (a[:lo] == -1).all() and (a[hi:] == -1).all()
        # gather: complete planes everywhere
        a = mine(truth)
        gather_rows([a], bounds)
        assert torch.equal(a, truth)
        # pyramid level gather: level k has max(h >> k, 1) rows; rank r owns rows y0 >> k .. y1 >> k (last strip: to the end)
        for k in (1, 3, 6):
            hk = max(h >> k, 1)
            lvl_truth = torch.arange(hk * 5, dtype=torch.float32).reshape(hk, 5)
            lvl = torch.full_like(lvl_truth, -1.0)
            lo = y0 >> k
            hi = hk if y1 == h else (y1 >> k)
            lvl[lo:hi] = lvl_truth[lo:hi]
            gather_rows([lvl], bounds, row_shift=k)
            assert torch.equal(lvl, lvl_truth), (rank, k)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 256), (3, 448), (2, 187)])
def test_exchange_over_gloo(world, h):
    mp.spawn(_worker, args=(world, _free_port(), h, 7), nprocs=world, join=True)


def test_rebalance_bounds_from_measured_times():
    """Measured feedback: the cuts move towards the rank that took longer, stay 64-row aligned, cover the frame, and a balanced
    measurement is a fixed point."""
    from diligentfx_b200.strips import rebalance_bounds
    b = [(0, 2624), (2624, 4320)]
    nb = rebalance_bounds(b, [3.08, 3.39], 4320)
    assert nb[0][0] == 0 and nb[-1][1] == 4320 and nb[0][1] == nb[1][0] and nb[0][1] % 64 == 0
    assert nb[0][1] > 2624                      # rank 1 was slower: it gives rows away
    assert rebalance_bounds(b, [3.2, 3.2 * 1696 / 2624 * 2624 / 1696], 4320)[0][1] in (2624, 2560, 2688)
    even = [(0, 1088), (1088, 2176), (2176, 3264), (3264, 4320)]
    assert rebalance_bounds(even, [1.0, 1.0, 1.0, 1.0 * 1056 / 1088], 4320) == even
    # an empty strip contributes nothing and may get rows back
    out = rebalance_bounds([(0, 64), (64, 64), (64, 432)], [1.0, 0.0, 5.0], 432)
    assert out[0][0] == 0 and out[-1][1] == 432 and all(a % 64 == 0 for a, _ in out)
