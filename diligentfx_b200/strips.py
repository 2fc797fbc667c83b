"""Row-strip sharding of one frame across the GPUs of a box (BASELINE.json config 4).

`SsrStrips` (bottom of the file) is the product path: the native executor of csrc/dfx_strips.cu (peer stores + flags for halo rows, peer
loads for the ray march and the temporal history, no NCCL call per frame). `strip_bounds` / `reflective_block_cost` decide who owns
which rows; `PeerSlab` allocates and exchanges the CUDA-IPC slabs once at start-up. `exchange_halo` / `gather_rows` are the generic
NCCL / gloo row exchanges of round 1 (torch.distributed P2P): kept as host-side utilities, no longer on the frame path.


Every rank holds FULL-SIZE planes and owns the rows `[y0, y1)` of every plane; the pass-level C-ABI computes only the owned
rows (`dfx_rows`). Between passes the host exchanges exactly the rows the next pass reads outside its strip:

* bounded reach  -> `exchange_halo`: grouped NCCL send/recv (`torch.distributed.batch_isend_irecv`) of `halo` rows with the
  upper and lower neighbour, written in place into the receiver's full-size plane;
* unbounded reach (SSR rays may cross the whole screen and fetch colour / normal at the hit; AO taps scale with 1/z)
  -> `gather_rows`: every rank sends its strip to every other rank (strips are 64-row aligned, hence unequal, so this is a
  grouped send/recv too rather than an equal-chunk all-gather).

Because every kernel addresses pixels by their global coordinates in full-size planes, a sharded run reads exactly the
values a single-GPU run reads: outputs are bit-identical (checked by tests/test_strips_gpu.py on 2 GPUs). The exchange
helpers are backend-agnostic (`gloo` on CPU tensors in tests/test_strips_cpu.py, `nccl` on device planes in production).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

STRIP_ALIGN = 64  # 2^SSR_DEPTH_HIERARCHY_MAX_MIP: pyramid passes then need no halo (SURVEY.md §8e "Partitioning")


def strip_bounds(height: int, world: int, align: int = STRIP_ALIGN, weights=None) -> list[tuple[int, int]]:
    """Rows [y0, y1) per rank: boundaries are multiples of `align`; the last strip ends at height.

    Without `weights` the strips have equal row counts (sizes differ by at most `align`). `weights` (one cost per `align`-row
    block, top to bottom) balances COST instead: SSR only marches rays for reflective pixels, which a typical frame
    concentrates in its lower half, so equal-height strips leave the top GPUs idle (see `reflective_block_cost`). Boundaries
    go where the running cost crosses k/world of the total; every rank keeps at least one block while blocks last.
    """
    blocks = -(-height // align)
    if weights is None:
        base, extra = divmod(blocks, world)
        counts = [base + (1 if r < extra else 0) for r in range(world)]
    else:
        w = [max(float(x), 0.0) for x in weights]
        assert len(w) == blocks, f"need one weight per {align}-row block ({blocks}), got {len(w)}"
        prefix = [0.0]
        for x in w:
            prefix.append(prefix[-1] + x)
        cuts: list[int] = []
        for r in range(1, world):
            target = prefix[-1] * r / world
            c = min(range(blocks + 1), key=lambda i: (abs(prefix[i] - target), i))  # cut closest to the target cost
            prev = cuts[-1] if cuts else 0
            if blocks >= world:
                c = min(max(c, prev + 1), blocks - (world - r))  # at least one block for every rank
            else:
                c = min(max(c, prev), blocks)
            cuts.append(c)
        edges = [0] + cuts + [blocks]
        counts = [edges[i + 1] - edges[i] for i in range(world)]
    bounds, y = [], 0
    for n in counts:
        y1 = min(height, y + n * align)
        bounds.append((y, y1))
        y = y1
    assert bounds[-1][1] == height
    return bounds


def reflective_block_cost(reflective_fraction_per_row, height: int, align: int = STRIP_ALIGN, march_cost: float = 11.0) -> list[float]:
    """Per-block cost for `strip_bounds(weights=…)` from the fraction of reflection samples in each row (any resolution: it is
    resampled to `height` rows). A pixel costs 1 (PostFX prep, Hi-Z, mask, the masked passes' early exit); a reflective one
    `march_cost` more (S4-S7; ratio measured on B200 at 4K: 1.3 ms over 66 % of the pixels against 0.175 ms over all)."""
    import numpy as np
    f = np.asarray(reflective_fraction_per_row, dtype=np.float64)
    rows = f[np.minimum((np.arange(height) * len(f)) // height, len(f) - 1)]
    cost = 1.0 + march_cost * rows
    blocks = -(-height // align)
    return [float(cost[b * align:min((b + 1) * align, height)].sum()) for b in range(blocks)]


def rebalance_bounds(bounds: list[tuple[int, int]], rank_ms: list[float], height: int, align: int = STRIP_ALIGN) -> list[tuple[int, int]]:
    """New strips from MEASURED per-rank compute times of the current ones: the time of a rank is spread evenly over its rows (a
    piecewise-constant cost density over the frame) and the cuts go where the cumulative cost crosses k / world of the total, at
    `align`-row granularity. A static model (`reflective_block_cost`) cannot know that rays starting on the ground plane march
    longer than rays starting on an object; two rounds of this feedback bring the ranks within a block of each other."""
    dens = [ms / max(b - a, 1) for ms, (a, b) in zip(rank_ms, bounds)]
    blocks = -(-height // align)
    weights = []
    for k in range(blocks):
        y0, y1 = k * align, min((k + 1) * align, height)
        w = 0.0
        for d, (a, b) in zip(dens, bounds):
            w += d * max(0, min(b, y1) - max(a, y0))
        weights.append(w)
    return strip_bounds(height, len(bounds), align, weights)


def exchange_halo(planes: list[torch.Tensor], bounds: list[tuple[int, int]], halo: int, group=None) -> None:
    """In place: after the call rows [y0 - halo, y0) and [y1, y1 + halo) of every plane hold the neighbours' owned rows.

    `planes` are full-size (H, W[, C]) tensors; this rank owns bounds[rank]. Ranks with an empty strip take no part.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1 or halo <= 0:
        return
    y0, y1 = bounds[rank]
    if y1 <= y0:
        return
    ops = []
    up = next((r for r in range(rank - 1, -1, -1) if bounds[r][1] > bounds[r][0]), None)
    down = next((r for r in range(rank + 1, world) if bounds[r][1] > bounds[r][0]), None)
    H = planes[0].shape[0]
    for p in planes:
        if up is not None:
            n_send = min(halo, y1 - y0)
            n_recv = min(halo, bounds[up][1] - bounds[up][0])
            ops.append(dist.P2POp(dist.isend, p[y0:y0 + n_send], up, group))
            ops.append(dist.P2POp(dist.irecv, p[y0 - n_recv:y0], up, group))
        if down is not None:
            n_send = min(halo, y1 - y0)
            n_recv = min(halo, bounds[down][1] - bounds[down][0], H - y1)
            ops.append(dist.P2POp(dist.isend, p[y1 - n_send:y1], down, group))
            ops.append(dist.P2POp(dist.irecv, p[y1:y1 + n_recv], down, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def gather_rows(planes: list[torch.Tensor], bounds: list[tuple[int, int]], group=None, row_shift: int = 0) -> None:
    """In place: after the call every plane is complete on every rank (each rank contributed its owned rows).

    `row_shift` = k gathers level k of a pyramid whose level-0 strips are `bounds` (rows y >> k; the last strip ends at the
    level's height).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    ops = []
    for p in planes:
        h = p.shape[0]

        def rows_of(r: int) -> tuple[int, int]:
            a, b = bounds[r]
            lo = a >> row_shift
            hi = h if r == world - 1 or bounds[r][1] >= bounds[-1][1] else (b >> row_shift)
            return lo, max(lo, min(hi, h))

        my = rows_of(rank)
        for r in range(world):
            if r == rank:
                continue
            theirs = rows_of(r)
            if my[1] > my[0]:
                ops.append(dist.P2POp(dist.isend, p[my[0]:my[1]], r, group))
            if theirs[1] > theirs[0]:
                ops.append(dist.P2POp(dist.irecv, p[theirs[0]:theirs[1]], r, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _RawCuda:
    """`__cuda_array_interface__` carrier: lets torch view device memory this package allocated itself."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerSlab:
    """One device allocation per rank, laid out identically on every rank and mapped into every peer (CUDA IPC, one process
    per GPU on one box), holding the planes other GPUs load from directly over NVLink.

    `specs`: name -> (shape, torch dtype). `self.local[name]` is this rank's tensor; `self.ptr(name, r)` is the address of
    rank r's copy as mapped into THIS process (usable by kernels launched on this rank's device). The mapping is opened
    with the reading device current (dfx_ipc_open), which is what makes it kernel-addressable; a mapping opened on the
    owner's device, as torch's tensor IPC does, serves copies but faults under direct kernel loads from another GPU.
    """

    ALIGN = 256

    def __init__(self, specs: dict, group=None, device: torch.device | None = None, fill: float = 0.0):
        from . import capi
        self.capi, self.lib, self.group = capi, capi.load(), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.offsets, off = {}, 0
        for name, (shape, dtype) in specs.items():
            n = int(torch.tensor(shape).prod().item()) * torch.empty((), dtype=dtype).element_size()
            self.offsets[name] = (off, n, tuple(shape), dtype)
            off += -(-n // self.ALIGN) * self.ALIGN
        self.nbytes = off
        handle = (C.c_uint8 * 64)()
        base = C.c_void_p()
        with torch.cuda.device(self.dev):
            capi.check(self.lib.dfx_ipc_alloc(C.c_size_t(self.nbytes), C.byref(base), handle), "dfx_ipc_alloc")
            self.base = [0] * self.world
            self.base[self.rank] = int(base.value)
            handles: list = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, bytes(handle), group=group)
            self._mapped = []
            for r in range(self.world):
                if r == self.rank:
                    continue
                ptr = C.c_void_p()
                capi.check(self.lib.dfx_ipc_open((C.c_uint8 * 64).from_buffer_copy(handles[r]), C.byref(ptr)), "dfx_ipc_open")
                self.base[r] = int(ptr.value)
                self._mapped.append(int(ptr.value))
        self._raw = _RawCuda(self.base[self.rank], self.nbytes)
        self._bytes = torch.as_tensor(self._raw, device=self.dev)
        assert self._bytes.data_ptr() == self.base[self.rank], "torch copied the slab instead of viewing it"
        self.local = {}
        for name, (o, n, shape, dtype) in self.offsets.items():
            self.local[name] = self._bytes[o:o + n].view(dtype).view(shape)
            self.local[name].fill_(fill)
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier(group)  # nobody reads a peer's slab before its owner has initialised it

    def ptr(self, name: str, r: int) -> int:
        return self.base[r] + self.offsets[name][0]

    def close(self):
        """Collective: unmap the peers' slabs, then free the own one once every peer has unmapped it."""
        if self.base is None:
            return
        torch.cuda.synchronize(self.dev)
        with torch.cuda.device(self.dev):
            for p in self._mapped:
                self.capi.check(self.lib.dfx_ipc_close(C.c_void_p(p)), "dfx_ipc_close")
            if self.world > 1:
                dist.barrier(self.group)
            self.local, self._bytes, self._raw = {}, None, None
            self.capi.check(self.lib.dfx_ipc_free(C.c_void_p(self.base[self.rank])), "dfx_ipc_free")
        self.base = None


class SsrStrips:
    """ScreenSpaceReflection (S1-S7 + the PostFX planes it reads) on one row strip of a frame whose other strips other ranks compute:
    a thin view of the native executor `dfx_ssr_strips_*` (csrc/dfx_strips.cu). All planes live in one slab per rank that every other
    rank has mapped; halo rows travel by peer stores + flags, the ray march and the temporal pass load from the owning GPU, and a frame
    costs this class ONE native call (no NCCL, no host synchronisation).

    Two ways to get the ranks together:
      * `SsrStrips.distributed(w, h, bounds, group)` - one process per GPU (torchrun): slabs are CUDA-IPC allocations exchanged over
        `torch.distributed` once at start-up;
      * `SsrStrips.virtual(w, h, bounds)` - all ranks in this process on the current device, one stream each (functional tests on
        a single GPU: the same kernels, flags and peer addressing, minus NVLink).
    """

    PLANES = {"depth": 0, "prev_depth": 1, "motion": 2, "normal": 3, "color": 4, "material": 5, "reproj": 6, "closest": 7, "previous_depth": 8,
              "roughness": 15, "mask": 16, "radiance": 17, "raydir": 18, "resolved_radiance": 19, "resolved_variance": 20, "resolved_depth": 21, "out": 26}
    INPUTS = ("depth", "prev_depth", "motion", "normal", "color", "material")

    def __init__(self, width: int, height: int, bounds, rank: int, bases: list[int], device=None):
        from . import capi
        self.capi, self.lib = capi, capi.load()
        self.w, self.h, self.rank, self.world, self.bounds = width, height, rank, len(bounds), list(bounds)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        assert self.bounds[0][0] == 0 and self.bounds[-1][1] == height and all(a % STRIP_ALIGN == 0 and a <= b for a, b in self.bounds)
        pm = capi.PeerMap()
        pm.count, pm.rank = self.world, rank
        for r, (a, _) in enumerate(self.bounds):
            pm.row_begin[r] = a
        pm.row_begin[self.world] = height
        for r in range(self.world):
            pm.base[r] = bases[r]
        self.handle = C.c_void_p()
        blob = open(capi.REPO_ROOT + "/diligentfx_b200/data/blue_noise_tables.bin", "rb").read()
        with torch.cuda.device(self.dev):
            capi.check(self.lib.dfx_ssr_strips_create(width, height, C.byref(pm), blob, C.byref(self.handle)), "dfx_ssr_strips_create")
        self.y0, self.y1 = self.bounds[rank]
        self._keep = None

    # ---- construction -------------------------------------------------------------------------------------------------------------
    @staticmethod
    def slab_bytes(width: int, height: int) -> int:
        from . import capi
        L = capi.load()
        L.dfx_ssr_strips_slab_bytes.restype = C.c_size_t
        return int(L.dfx_ssr_strips_slab_bytes(width, height))

    @classmethod
    def virtual(cls, width: int, height: int, bounds, device=None) -> list["SsrStrips"]:
        dev = device or torch.device("cuda", torch.cuda.current_device())
        n = cls.slab_bytes(width, height)
        slabs = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in bounds]
        ranks = [cls(width, height, bounds, r, [t.data_ptr() for t in slabs], dev) for r in range(len(bounds))]
        for r, x in enumerate(ranks):
            x._keep = slabs
            x.stream = torch.cuda.Stream(dev)
        torch.cuda.synchronize(dev)
        return ranks

    @classmethod
    def distributed(cls, width: int, height: int, bounds, group=None, device=None, slab: "PeerSlab | None" = None) -> "SsrStrips":
        """`slab`: reuse the CUDA-IPC slab of an executor that was closed with `close(keep_slab=True)` (new strips, same memory)."""
        if dist.get_world_size(group) > 1:
            dist.barrier(group)  # nobody is still pushing into a slab that is about to be zeroed
        slab = slab or PeerSlab({"slab": ((cls.slab_bytes(width, height),), torch.uint8)}, group, device)
        x = cls(width, height, bounds, dist.get_rank(group), list(slab.base), device)
        x._keep = slab
        x.stream = torch.cuda.current_stream(x.dev)
        torch.cuda.synchronize(x.dev)
        if dist.get_world_size(group) > 1:
            dist.barrier(group)  # every rank has zeroed its slab before anybody pushes into it
        return x

    def close(self, keep_slab: bool = False):
        """Collective when the slab is a CUDA-IPC one. `keep_slab=True` returns it instead of freeing it (see `distributed(slab=...)`)."""
        if self.handle:
            torch.cuda.synchronize(self.dev)
            self.lib.dfx_ssr_strips_destroy(self.handle)
            self.handle = None
        kept = self._keep
        self._keep = None
        if isinstance(kept, PeerSlab) and not keep_slab:
            kept.close()
            kept = None
        return kept

    # ---- planes -------------------------------------------------------------------------------------------------------------------
    def plane(self, name_or_id) -> "object":
        p = self.capi.Plane()
        self.capi.check(self.lib.dfx_ssr_strips_plane(self.handle, self.PLANES.get(name_or_id, name_or_id), C.byref(p)), "dfx_ssr_strips_plane")
        return p

    def _rows_of(self, p, y0: int, y1: int):
        q = self.capi.Plane(p.ptr + y0 * p.pitch_bytes, p.pitch_bytes, p.width, y1 - y0, p.format, p.flags)
        return q

    def write_inputs(self, frame: dict, rows: tuple[int, int] | None = None, stream=None):
        """Host planes (numpy, full frame) -> this rank's slab, the owned rows only (what a renderer that shards its G-buffer pass the
        same way would produce in place)."""
        import numpy as np
        y0, y1 = rows or (self.y0, self.y1)
        s = C.c_void_p((stream or torch.cuda.current_stream(self.dev)).cuda_stream)
        for name in self.INPUTS:
            a = np.ascontiguousarray(frame[name][y0:y1], np.float32)
            q = self._rows_of(self.plane(name), y0, y1)
            self.capi.check(self.lib.dfx_plane_upload(s, C.byref(q), a.ctypes.data_as(C.c_void_p), 0), "dfx_plane_upload")
        self.capi.check(self.lib.dfx_stream_synchronize(s))  # pageable source

    def read(self, name: str, rows: tuple[int, int] | None = None):
        import numpy as np
        y0, y1 = rows or (self.y0, self.y1)
        p = self.plane(name)
        ch = {self.capi.FORMAT_R32F: 1, self.capi.FORMAT_RG32F: 2, self.capi.FORMAT_RGBA32F: 4}[p.format]
        out = np.empty((y1 - y0, p.width) if ch == 1 else (y1 - y0, p.width, ch), np.float32)
        torch.cuda.synchronize(self.dev)
        q = self._rows_of(p, y0, y1)
        self.capi.check(self.lib.dfx_plane_download(None, C.byref(q), out.ctypes.data_as(C.c_void_p), 0), "dfx_plane_download")
        self.capi.check(self.lib.dfx_stream_synchronize(None))
        return out

    # ---- one frame ----------------------------------------------------------------------------------------------------------------
    def execute(self, frame_index: int, curr_camera, prev_camera, attribs=None, stream=None):
        a = attribs or self.capi.SSRAttribs.default()
        s = C.c_void_p((stream or getattr(self, "stream", None) or torch.cuda.current_stream(self.dev)).cuda_stream)
        self.capi.check(self.lib.dfx_ssr_strips_execute(self.handle, s, frame_index, C.byref(curr_camera), C.byref(prev_camera), C.byref(a)), "dfx_ssr_strips_execute")

    def timed_out(self) -> bool:
        t = C.c_int32()
        self.capi.check(self.lib.dfx_ssr_strips_check(self.handle, C.byref(t)))
        return bool(t.value)
