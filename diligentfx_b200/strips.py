"""Row-strip sharding of one frame across the GPUs of a box (BASELINE.json config 4).

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


class SsrStripRunner:
    """ScreenSpaceReflection (S1-S7) + the PostFX planes it needs, one frame split into row strips over the ranks of `group`.

    Reach of every pass: P1-P3 +-1 row of depth / motion; S1, S2 none (64-row aligned strips); S4 unbounded
    (Hi-Z, colour, normal gathered); S5 +-4 rows of the intersect outputs; S6 previous-frame planes at the reprojected
    position (|motion| is capped at MAX_MOTION_ROWS by the caller) plus +-1 row of the resolved radiance; S7 +-2 rows.
    """

    MAX_MOTION_ROWS = 24  # reprojection reach (motion + 3x3 search + bilinear footprint) the temporal pass is given

    def __init__(self, width: int, height: int, group=None, device: torch.device | None = None, peer: bool = False, poison: bool = False,
                 input_sets: int = 1, bounds: list[tuple[int, int]] | None = None):
        """`peer=True`: no gather before the ray march — the intersect kernel loads Hi-Z / colour / normal texels straight from
        the GPU that owns their row over NVLink (dfx_pass_ssr_intersect_peer). The runner then owns the depth / colour /
        normal planes the peers read: `self.shared_sets[i]` for i < input_sets (a renderer that double-buffers its G-buffer
        asks for 2). Fill a set directly and name it in execute(input_set=i), or pass other tensors to execute() and pay a
        device copy of the strip. `poison=True` fills those planes with NaN first (tests: a texel read from a row nobody
        wrote shows up in the output). `bounds`: explicit strips (e.g. cost-balanced, `strip_bounds(weights=…)`)."""
        from . import capi
        self.capi = capi
        self.lib = capi.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.w, self.h = width, height
        self.bounds = list(bounds) if bounds is not None else strip_bounds(height, self.world)  # identical on every rank
        assert len(self.bounds) == self.world and self.bounds[0][0] == 0 and self.bounds[-1][1] == height
        assert all(a % STRIP_ALIGN == 0 and a <= b for a, b in self.bounds), "strip boundaries must be multiples of 64 rows"
        self.rows = capi.Rows(*self.bounds[self.rank])
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        H, W = height, width
        self.hiz = [None] + [f(max(H >> i, 1), max(W >> i, 1)) for i in range(1, 7)]
        self.roughness, self.mask = f(H, W), torch.zeros((H, W), dtype=torch.uint8, device=dev)
        self.radiance, self.raydir = f(H, W, 4), f(H, W, 4)
        self.res_rad, self.res_var, self.res_depth = f(H, W, 4), f(H, W), f(H, W)
        self.radhist, self.varhist = [f(H, W, 4), f(H, W, 4)], [f(H, W), f(H, W)]
        self.out = f(H, W, 4)
        self.reproj, self.closest, self.prev_depth = f(H, W), f(H, W, 2), f(H, W)
        self.bn_xy, self.bn_zw = f(128, 128, 2), f(128, 128, 2)
        blob = open(capi.REPO_ROOT + "/diligentfx_b200/data/blue_noise_tables.bin", "rb").read()
        self.tables = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        self.cams = torch.zeros(2 * 576, dtype=torch.uint8, device=dev)
        self.comm_bytes = 0
        self.peer = bool(peer) and self.world > 1
        if self.peer:
            fill = float("nan") if poison else 0.0
            specs = {f"hiz{k}": (tuple(self.hiz[k].shape), torch.float32) for k in range(1, 7)}
            for i in range(max(1, input_sets)):
                specs.update({f"depth{i}": ((H, W), torch.float32), f"color{i}": ((H, W, 4), torch.float32), f"normal{i}": ((H, W, 4), torch.float32)})
            self.slab = PeerSlab(specs, group, dev, fill)
            self.hiz = [None] + [self.slab.local[f"hiz{k}"] for k in range(1, 7)]
            self.shared_sets = [{n: self.slab.local[f"{n}{i}"] for n in ("depth", "color", "normal")} for i in range(max(1, input_sets))]
            self.peer_sets = []
            for i in range(len(self.shared_sets)):
                ps = capi.PeerSet()
                ps.count = self.world
                for r, (a, _) in enumerate(self.bounds):
                    ps.row_begin[r] = a
                ps.row_begin[self.world] = H
                for r in range(self.world):
                    ps.color[r], ps.normal[r] = self.slab.ptr(f"color{i}", r), self.slab.ptr(f"normal{i}", r)
                    ps.hiz[0][r] = self.slab.ptr(f"depth{i}", r)
                    for k in range(1, 7):
                        ps.hiz[k][r] = self.slab.ptr(f"hiz{k}", r)
                self.peer_sets.append(ps)
            self.token = torch.zeros(1, dtype=torch.float32, device=dev)

    def close(self):
        """Collective in peer mode (unmaps / frees the shared slab)."""
        if self.peer:
            self.slab.close()
            self.peer = False

    def _device_barrier(self):
        """Stream-ordered barrier over the ranks (a 4-byte all-reduce): work enqueued after it on any rank starts only when
        the work enqueued before it on every rank has finished. Does not block the host."""
        dist.all_reduce(self.token, group=self.group)

    def _count(self, planes, rows: int):
        self.comm_bytes += sum(rows * p[0].numel() * p.element_size() for p in planes)

    def execute(self, frame_index: int, inputs: dict, curr_camera, prev_camera, attribs=None, flags: int = 0,
                input_set: int = 0) -> torch.Tensor:
        """`inputs`: full-size device planes depth, prev_depth, motion, normal, color, material of which this rank's strip is
        valid (everything else is filled in by the exchanges). Returns the SSR output plane (valid on the owned rows).
        Peer mode: depth / color / normal are taken from `self.shared_sets[input_set]`; entries of `inputs` under those
        names that are other tensors are first copied into the set (owned rows)."""
        capi, L, B, R, g = self.capi, self.lib, self.bounds, self.rows, self.group
        P = capi.plane_of
        a = attribs or capi.SSRAttribs.default()
        s = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        cur, prv = frame_index & 1, (frame_index + 1) & 1
        # cameras: pinned ring + async copy. A blocking copy here would make the host wait for the previous frame's kernels every
        # frame, i.e. serialise launch overhead with execution (measured: the difference between 1.1x and real strong scaling).
        if not hasattr(self, "_cam_ring"):
            self._cam_ring = [(torch.empty(2 * 576, dtype=torch.uint8).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._cam_next = 0
        buf, ev = self._cam_ring[self._cam_next]
        self._cam_next = (self._cam_next + 1) % len(self._cam_ring)
        ev.synchronize()  # the copy issued from this slot four frames ago has executed
        buf.copy_(torch.frombuffer(bytearray(bytes(curr_camera) + bytes(prev_camera)), dtype=torch.uint8))
        self.cams.copy_(buf, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.dev))
        cams = C.c_void_p(self.cams.data_ptr())
        ck = capi.check
        depth, motion, normal, color = inputs.get("depth"), inputs["motion"], inputs.get("normal"), inputs.get("color")
        if self.peer:
            if flags & capi.SSR_FLAG_PREVIOUS_FRAME:
                raise capi.DfxError("previous-frame SSR is not supported on peer-sharded frames")
            shared = self.shared_sets[input_set]
            for name in ("depth", "color", "normal"):
                if inputs.get(name) is not None and inputs[name] is not shared[name]:
                    shared[name][R.y0:R.y1].copy_(inputs[name][R.y0:R.y1])
            depth, normal, color = shared["depth"], shared["normal"], shared["color"]

        # PostFX prep: 3x3 closest-depth search -> +-1 row of depth and motion; previous depth: reprojection reach
        # (peer mode: S5 / S7 read depth and normal of up to 4 rows beyond the strip, which the gather would have provided)
        if self.peer:
            exchange_halo([depth, normal], B, 4, g)
            exchange_halo([motion], B, 1, g)
        else:
            exchange_halo([depth, motion], B, 1, g)
        exchange_halo([inputs["prev_depth"]], B, self.MAX_MOTION_ROWS, g)
        ck(L.dfx_pass_blue_noise(s, C.c_void_p(self.tables.data_ptr()), frame_index, C.byref(P(self.bn_xy)), C.byref(P(self.bn_zw))))
        # the copy of the previous depth has to cover the halo rows the temporal pass will read: widen the row range
        y0w, y1w = max(R.y0 - self.MAX_MOTION_ROWS, 0), min(R.y1 + self.MAX_MOTION_ROWS, self.h)
        ck(L.dfx_pass_postfx_prepare(s, cams, C.byref(P(depth)), C.byref(P(inputs["prev_depth"])), C.byref(P(motion)), C.byref(P(self.reproj)),
                                     C.byref(P(self.closest)), C.byref(P(self.prev_depth)), R))
        self.prev_depth[y0w:R.y0].copy_(inputs["prev_depth"][y0w:R.y0])
        self.prev_depth[R.y1:y1w].copy_(inputs["prev_depth"][R.y1:y1w])

        # S1 + S2 on the owned rows, then make Hi-Z / colour / normal complete everywhere for the ray march
        pyr = capi.pyramid_of([depth] + self.hiz[1:])
        ck(L.dfx_pass_ssr_hiz(s, C.byref(pyr), R))
        ck(L.dfx_pass_ssr_mask_roughness(s, C.byref(a), C.byref(P(inputs["material"])), C.byref(P(depth)), C.byref(P(self.roughness)), C.byref(P(self.mask)), R))
        if self.peer:
            # S4 with peer loads: every rank's Hi-Z / colour / normal strips must be complete before anybody marches, and
            # every march must be over before anybody overwrites them (next frame) -> one device barrier on either side
            self._device_barrier()
            ck(L.dfx_pass_ssr_intersect_peer(s, cams, C.byref(a), flags, C.byref(self.peer_sets[input_set]), C.byref(P(color)), C.byref(P(normal)),
                                             C.byref(P(self.roughness)), C.byref(P(self.mask)), C.byref(P(self.bn_xy)), C.byref(pyr),
                                             C.byref(P(self.radiance)), C.byref(P(self.raydir)), R))
            self._device_barrier()
        else:
            gather_rows([depth, color, normal] + ([motion] if flags & capi.SSR_FLAG_PREVIOUS_FRAME else []), B, g)
            for k in range(1, 7):
                gather_rows([self.hiz[k]], B, g, row_shift=k)
            # S4
            ck(L.dfx_pass_ssr_intersect(s, cams, C.byref(a), flags, C.byref(P(color)), C.byref(P(normal)), C.byref(P(self.roughness)), C.byref(P(self.mask)),
                                        C.byref(P(self.bn_xy)), C.byref(pyr), C.byref(P(motion)), C.byref(P(self.radiance)), C.byref(P(self.raydir)), R))
        # S5: 8-tap disk of radius <= 4 px
        exchange_halo([self.radiance, self.raydir], B, 4, g)
        ck(L.dfx_pass_ssr_spatial(s, cams, C.byref(a), C.byref(P(self.roughness)), C.byref(P(self.mask)), C.byref(P(normal)), C.byref(P(depth)),
                                  C.byref(P(self.raydir)), C.byref(P(self.radiance)), C.byref(P(self.res_rad)), C.byref(P(self.res_var)),
                                  C.byref(P(self.res_depth)), R))
        # S6: 3x3 statistics of the resolved radiance; history of the previous frame at the reprojected position
        exchange_halo([self.res_rad], B, 1, g)
        exchange_halo([self.radhist[prv], self.varhist[prv]], B, self.MAX_MOTION_ROWS, g)
        ck(L.dfx_pass_ssr_temporal(s, cams, C.byref(a), C.byref(P(self.mask)), C.byref(P(motion)), C.byref(P(self.res_depth)), C.byref(P(self.reproj)),
                                   C.byref(P(self.res_rad)), C.byref(P(self.res_var)), C.byref(P(self.prev_depth)), C.byref(P(self.radhist[prv])),
                                   C.byref(P(self.varhist[prv])), C.byref(P(self.radhist[cur])), C.byref(P(self.varhist[cur])), R))
        # S7: (2r+1)^2 window, r <= 2, reads roughness / radiance of the neighbours (normal and depth are already complete)
        exchange_halo([self.radhist[cur], self.roughness], B, 2, g)
        ck(L.dfx_pass_ssr_bilateral(s, cams, C.byref(a), C.byref(P(self.mask)), C.byref(P(depth)), C.byref(P(normal)), C.byref(P(self.roughness)),
                                    C.byref(P(self.radhist[cur])), C.byref(P(self.varhist[cur])), C.byref(P(self.out)), R))
        return self.out
