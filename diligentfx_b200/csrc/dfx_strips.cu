// dfx_strips.cu — one frame split into row strips over the GPUs of a box (BASELINE.json config 4: ScreenSpaceReflection
// on a 7680x4320 G-buffer at 2 / 4 / 8 GPUs). No reference counterpart: the reference renders a frame on one device.
//
// Layout: every rank holds ALL planes of the effect, full size, at the same offsets of an identically laid-out slab that every
// other rank has mapped (CUDA IPC, one process per GPU; or plain pointers when several "ranks" share a process). Rank r owns the
// 64-row aligned strip [row_begin[r], row_begin[r+1]) of every plane and computes only those rows (the pass-level C-ABI takes a
// row range). What a pass reads outside its strip arrives in one of two ways, both without the host and without NCCL:
//   * bounded taps (3x3, +-4 px disk, 5x5 window)  -> HALO PUSH: the producer copies its first / last few rows into the neighbour's
//     slab with plain stores over NVLink (halo_push_kernel) and then raises a flag in the neighbour's slab; the consumer's stream
//     holds a one-thread kernel that spins on that flag (flag_wait_kernel). Pure dataflow between neighbours: nobody waits for a
//     rank it does not read from, and the transfer overlaps whatever else the GPUs are doing;
//   * unbounded reach (the Hi-Z ray march may cross the whole frame and fetch colour / normal at the hit; the temporal pass reads
//     last frame's planes at reprojected positions)  -> PEER LOADS: the kernel itself loads each texel from the slab of the GPU that
//     owns its row (ssr_intersect_kernel<PEER>, ssr_temporal_kernel<PEER>), behind an all-rank barrier of the same flag kind.
// Every kernel addresses texels by their global coordinates in full-size planes, so a sharded frame reads exactly the values the
// unsharded frame reads: outputs are bit-identical (tests/test_strips_gpu.py, bench.py's strips leg).
#include "dfx_common.cuh"
#include <algorithm>
#include <new>

using namespace dfx;

namespace
{
constexpr int      kMaxPushPlanes = 6;
constexpr int      kFlagSlots     = 16;
constexpr unsigned kSpinLimit     = 4000000000u; // ~2 s of SM clocks: a peer that never signals turns into an error flag, not a hung GPU

// ---- device-side synchronisation ------------------------------------------------------------------------------------------------
struct SyncBlock // lives at the head of every slab
{
    unsigned from_up[kFlagSlots], from_down[kFlagSlots]; // written by the upper / lower neighbour: "my halo rows for exchange k of frame seq are in your slab"
    unsigned barrier[DFX_MAX_PEERS];                     // barrier[r] written by rank r
    unsigned push_tickets[kFlagSlots];                   // local: blocks of a halo push that have finished
    unsigned error;                                      // a wait timed out
    unsigned pad[7];
};

__device__ __forceinline__ unsigned ld_volatile(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ bool     reached(unsigned flag, unsigned value) { return int(flag - value) >= 0; }
__device__ void spin_until(const unsigned* flag, unsigned value, unsigned* error)
{
    const long long t0 = clock64();
    while (!reached(ld_volatile(flag), value))
    {
        __nanosleep(64);
        if ((unsigned long long)(clock64() - t0) > kSpinLimit)
        {
            atomicExch(error, 1u);
            break;
        }
    }
}

struct PushSeg
{
    const char* src;    // first row of the range in this rank's plane
    long long   delta;  // byte distance to the same address in the neighbour's slab
    long long   pitch;  // bytes between rows
    int         rows, row_chunks; // 16-byte chunks per row
};
constexpr int kMaxPushSegs = 48; // halo pushes use up to 2 x kMaxPushPlanes; the Hi-Z all-gather 6 levels x 7 peers
struct PushArgs
{
    PushSeg   seg[kMaxPushSegs];
    int       nseg;
    unsigned* flag_up;   // in the upper neighbour's slab (nullptr: no such neighbour)
    unsigned* flag_down; // in the lower neighbour's slab
    unsigned* tickets;   // local
    unsigned  value;
};
// Copies the listed row ranges into the neighbours' slabs (128-bit stores to peer memory), then the last block to finish raises the flags.
__global__ void __launch_bounds__(256) halo_push_kernel(const __grid_constant__ PushArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int s = 0; s < a.nseg; ++s)
    {
        const PushSeg& g = a.seg[s];
        const int      n = g.rows * g.row_chunks;
        for (int i0 = tid; i0 < n; i0 += 4 * nth) // four 16-byte chunks in flight per thread: the link's latency is long
        {
            float4      v[4];
            const char* p[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const int i = i0 + k * nth;
                if (i < n)
                {
                    const int r = i / g.row_chunks, c = i - r * g.row_chunks;
                    p[k]        = g.src + r * g.pitch + (long long)c * 16;
                    v[k]        = __ldg(reinterpret_cast<const float4*>(p[k]));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i0 + k * nth < n) *reinterpret_cast<float4*>(const_cast<char*>(p[k]) + g.delta) = v[k];
        }
    }
    __threadfence_system(); // this thread's stores are visible system-wide before the ticket below
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(a.tickets, 1u);
        if (t == gridDim.x - 1)
        {
            *a.tickets = 0;
            __threadfence_system();
            if (a.flag_up) *reinterpret_cast<volatile unsigned*>(a.flag_up) = a.value;
            if (a.flag_down) *reinterpret_cast<volatile unsigned*>(a.flag_down) = a.value;
        }
    }
}
__global__ void flag_wait_kernel(const unsigned* from_up, const unsigned* from_down, unsigned value, unsigned* error)
{
    if (from_up) spin_until(from_up, value, error);
    if (from_down) spin_until(from_down, value, error);
    __threadfence_system();
}
struct BarrierArgs
{
    unsigned* remote[DFX_MAX_PEERS]; // &barrier[me] in rank r's slab
    unsigned* local;                 // barrier[] in the own slab
    unsigned* error;
    int       count, me;
    unsigned  value;
};
// All-rank barrier in stream order: work enqueued after it on any rank starts only when the work enqueued before it on every rank has finished.
__global__ void barrier_kernel(const __grid_constant__ BarrierArgs a)
{
    const int r = threadIdx.x;
    if (r < a.count && r != a.me)
    {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned*>(a.remote[r]) = a.value;
        spin_until(a.local + r, a.value, a.error);
    }
    __threadfence_system();
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
size_t texel_bytes(int fmt) { return fmt == DFX_FORMAT_R32F ? 4 : fmt == DFX_FORMAT_RG32F ? 8 : fmt == DFX_FORMAT_RGBA32F ? 16 : 1; }

} // namespace
namespace dfx
{
dfx_status launch_upload_cameras(void* stream, const dfx_camera_attribs* curr, const dfx_camera_attribs* prev, uint32_t frame_index, dfx_camera_attribs* dst_cams,
                                 uint32_t* dst_frame);
void       preload_ssr_strip_kernels(); // dfx_ssr.cu
}

struct dfx_ssr_strips
{
    int          w = 0, h = 0, world = 1, rank = 0, up = -1, down = -1;
    dfx_peer_map peers{};
    dfx_rows     rows{};
    char*        base = nullptr;
    dfx_plane    plane[DFX_SSR_STRIPS_PLANE_COUNT]{};
    dfx_plane    bn_xy{}, bn_zw{};
    dfx_camera_attribs* cams_dev = nullptr;
    uint8_t*     tables_dev = nullptr;
    SyncBlock*   sync = nullptr;
    unsigned     seq = 0;
    dfx_peer_set peer_set{};
    cudaStream_t side = nullptr;                 // the depth plane's all-gather runs beside the pre-march passes
    cudaEvent_t  ev_begin = nullptr, ev_side = nullptr;

    template <class T> T* remote(T* local_ptr, int r) const
    {
        return reinterpret_cast<T*>(reinterpret_cast<char*>(local_ptr) - base + const_cast<char*>(static_cast<const char*>(peers.base[r])));
    }
    ~dfx_ssr_strips()
    {
        if (side) cudaStreamDestroy(side);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_side) cudaEventDestroy(ev_side);
        cudaFree(tables_dev);
    }
};

// ---- slab layout: a pure function of the frame size ---------------------------------------------------------------------------------
namespace
{
struct Layout
{
    size_t offset[DFX_SSR_STRIPS_PLANE_COUNT], pitch[DFX_SSR_STRIPS_PLANE_COUNT];
    int    width[DFX_SSR_STRIPS_PLANE_COUNT], height[DFX_SSR_STRIPS_PLANE_COUNT], format[DFX_SSR_STRIPS_PLANE_COUNT];
    size_t sync, cams, bn_xy, bn_zw, total;
};
int plane_format(int id)
{
    switch (id)
    {
        case DFX_SSR_STRIPS_PLANE_MOTION: case DFX_SSR_STRIPS_PLANE_CLOSEST_MOTION: return DFX_FORMAT_RG32F;
        case DFX_SSR_STRIPS_PLANE_NORMAL: case DFX_SSR_STRIPS_PLANE_COLOR: case DFX_SSR_STRIPS_PLANE_MATERIAL: case DFX_SSR_STRIPS_PLANE_RADIANCE:
        case DFX_SSR_STRIPS_PLANE_RAYDIR: case DFX_SSR_STRIPS_PLANE_RESOLVED_RADIANCE: case DFX_SSR_STRIPS_PLANE_RADIANCE_HISTORY0:
        case DFX_SSR_STRIPS_PLANE_RADIANCE_HISTORY1: case DFX_SSR_STRIPS_PLANE_OUTPUT: return DFX_FORMAT_RGBA32F;
        case DFX_SSR_STRIPS_PLANE_MASK: return DFX_FORMAT_R8U;
        default: return DFX_FORMAT_R32F;
    }
}
Layout make_layout(int w, int h)
{
    Layout L{};
    size_t off = 0;
    L.sync     = off, off += align_up(sizeof(SyncBlock), 512);
    L.cams     = off, off += align_up(2 * sizeof(dfx_camera_attribs), 512);
    L.bn_xy    = off, off += 128 * 128 * 8;
    L.bn_zw    = off, off += 128 * 128 * 8;
    for (int id = 0; id < DFX_SSR_STRIPS_PLANE_COUNT; ++id)
    {
        const int mip = (id >= DFX_SSR_STRIPS_PLANE_HIZ1 && id < DFX_SSR_STRIPS_PLANE_HIZ1 + 6) ? id - DFX_SSR_STRIPS_PLANE_HIZ1 + 1 : 0;
        L.width[id] = std::max(w >> mip, 1), L.height[id] = std::max(h >> mip, 1), L.format[id] = plane_format(id);
        L.pitch[id]  = align_up(size_t(L.width[id]) * texel_bytes(L.format[id]), 128);
        L.offset[id] = off;
        off += align_up(L.pitch[id] * size_t(L.height[id]), 512);
    }
    L.total = off;
    return L;
}
} // namespace

extern "C" size_t dfx_ssr_strips_slab_bytes(int32_t width, int32_t height) { return (width > 0 && height > 0) ? make_layout(width, height).total : 0; }

extern "C" dfx_status dfx_ssr_strips_create(int32_t width, int32_t height, const dfx_peer_map* peers, const uint8_t* blue_noise_tables, dfx_ssr_strips** out)
{
    DFX_REQUIRE(out && peers && blue_noise_tables && width > 0 && height > 0, "bad arguments");
    PeerMap check;
    DFX_REQUIRE(make_peer_map(peers, height, check), "bad peer map (rank count, 64-row aligned strips, slab bases)");
    dfx_ssr_strips* s = new (std::nothrow) dfx_ssr_strips;
    DFX_REQUIRE(s, "out of memory");
    s->w = width, s->h = height, s->world = peers->count, s->rank = peers->rank, s->peers = *peers;
    s->rows = dfx_rows{peers->row_begin[s->rank], peers->row_begin[s->rank + 1]};
    for (int r = s->rank - 1; r >= 0 && s->up < 0; --r)
        if (peers->row_begin[r + 1] > peers->row_begin[r]) s->up = r;
    for (int r = s->rank + 1; r < s->world && s->down < 0; ++r)
        if (peers->row_begin[r + 1] > peers->row_begin[r]) s->down = r;
    s->base        = const_cast<char*>(static_cast<const char*>(peers->base[s->rank]));
    const Layout L = make_layout(width, height);
    for (int id = 0; id < DFX_SSR_STRIPS_PLANE_COUNT; ++id)
        s->plane[id] = dfx_plane{s->base + L.offset[id], L.pitch[id], L.width[id], L.height[id], L.format[id], 0};
    s->bn_xy    = dfx_plane{s->base + L.bn_xy, 128 * 8, 128, 128, DFX_FORMAT_RG32F, 0};
    s->bn_zw    = dfx_plane{s->base + L.bn_zw, 128 * 8, 128, 128, DFX_FORMAT_RG32F, 0};
    s->cams_dev = reinterpret_cast<dfx_camera_attribs*>(s->base + L.cams);
    s->sync     = reinterpret_cast<SyncBlock*>(s->base + L.sync);
    cudaError_t e = cudaMemset(s->base, 0, L.total); // histories start from 0 (ScreenSpaceReflection.cpp:263-264, :279-280), flags from 0
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->tables_dev, 256 + 128 * 128 * 8);
    if (e == cudaSuccess) e = cudaMemcpy(s->tables_dev, blue_noise_tables, 256 + 128 * 128 * 8, cudaMemcpyHostToDevice);
    // CUDA loads kernels lazily, and loading one may synchronise the device: with a flag-wait kernel already spinning on a stream that
    // would deadlock ranks sharing a process (and serialise the first frame otherwise). Everything a frame launches is loaded here.
    preload_ssr_strip_kernels();
    {
        cudaFuncAttributes fa;
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, halo_push_kernel);
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, flag_wait_kernel);
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, barrier_kernel);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_begin, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_side, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
        delete s;
        return check_cuda(e, "dfx_ssr_strips_create");
    }
    // the form dfx_pass_ssr_intersect_peer takes: per-plane per-owner pointers
    dfx_peer_set& ps = s->peer_set;
    ps.count         = s->world;
    for (int r = 0; r <= s->world; ++r) ps.row_begin[r] = peers->row_begin[r];
    for (int r = 0; r < s->world; ++r)
    {
        ps.color[r]  = s->remote(static_cast<char*>(s->plane[DFX_SSR_STRIPS_PLANE_COLOR].ptr), r);
        ps.normal[r] = s->remote(static_cast<char*>(s->plane[DFX_SSR_STRIPS_PLANE_NORMAL].ptr), r);
        ps.hiz[0][r] = s->remote(static_cast<char*>(s->plane[DFX_SSR_STRIPS_PLANE_DEPTH].ptr), r);
        for (int k = 1; k <= 6; ++k) ps.hiz[k][r] = s->plane[DFX_SSR_STRIPS_PLANE_HIZ1 + k - 1].ptr; // gathered before the march: every row is local
    }
    *out = s;
    return DFX_OK;
}
extern "C" void dfx_ssr_strips_destroy(dfx_ssr_strips* s) { delete s; }

extern "C" dfx_status dfx_ssr_strips_plane(const dfx_ssr_strips* s, int32_t id, dfx_plane* out)
{
    DFX_REQUIRE(s && out && id >= 0 && id < DFX_SSR_STRIPS_PLANE_COUNT, "bad plane id %d", id);
    *out = s->plane[id];
    return DFX_OK;
}
extern "C" dfx_status dfx_ssr_strips_rows(const dfx_ssr_strips* s, dfx_rows* out)
{
    DFX_REQUIRE(s && out, "null argument");
    *out = s->rows;
    return DFX_OK;
}
// 1 if a flag wait of an earlier frame timed out (a peer never signalled). Synchronises the device.
extern "C" dfx_status dfx_ssr_strips_check(const dfx_ssr_strips* s, int32_t* timed_out)
{
    DFX_REQUIRE(s && timed_out, "null argument");
    unsigned e = 0;
    DFX_CUDA(cudaMemcpy(&e, &s->sync->error, sizeof(e), cudaMemcpyDeviceToHost));
    *timed_out = int32_t(e);
    return DFX_OK;
}

namespace
{
struct PushSpec
{
    int id, up_rows, down_rows; // rows of the strip's top pushed to the upper neighbour / of its bottom pushed to the lower one
};
bool has_neighbours(const dfx_ssr_strips* s) { return s->world > 1 && (s->up >= 0 || s->down >= 0) && s->rows.y1 > s->rows.y0; }
unsigned flag_value(const dfx_ssr_strips* s, int slot) { return s->seq * kFlagSlots + unsigned(slot); }

// my halo rows -> the neighbours' slabs, then their flags
dfx_status push_halos(dfx_ssr_strips* s, cudaStream_t st, int slot, const PushSpec* specs, int n)
{
    if (!has_neighbours(s)) return DFX_OK;
    DFX_PROFILE(st, "strips_halo_push");
    PushArgs a{};
    const int strip = s->rows.y1 - s->rows.y0;
    for (int i = 0; i < n; ++i)
    {
        const dfx_plane& p = s->plane[specs[i].id];
        const int        chunks = int(align_up(size_t(p.width) * texel_bytes(p.format), 16) / 16);
        const char*      b = static_cast<const char*>(p.ptr);
        if (s->up >= 0 && specs[i].up_rows > 0)
        {
            const int rows = std::min(specs[i].up_rows, strip);
            a.seg[a.nseg++] = PushSeg{b + size_t(s->rows.y0) * p.pitch_bytes, s->remote(const_cast<char*>(b), s->up) - b, (long long)p.pitch_bytes, rows, chunks};
        }
        if (s->down >= 0 && specs[i].down_rows > 0)
        {
            const int rows = std::min(specs[i].down_rows, strip);
            a.seg[a.nseg++] = PushSeg{b + size_t(s->rows.y1 - rows) * p.pitch_bytes, s->remote(const_cast<char*>(b), s->down) - b, (long long)p.pitch_bytes, rows, chunks};
        }
    }
    a.value     = flag_value(s, slot);
    a.flag_up   = s->up >= 0 ? &s->remote(s->sync, s->up)->from_down[slot] : nullptr;     // I am my upper neighbour's lower neighbour
    a.flag_down = s->down >= 0 ? &s->remote(s->sync, s->down)->from_up[slot] : nullptr;
    a.tickets   = &s->sync->push_tickets[slot];
    halo_push_kernel<<<32, 256, 0, st>>>(a);
    DFX_LAUNCHED("halo_push_kernel");
    return DFX_OK;
}
// the neighbours' halo rows of exchange `slot` have arrived in my slab (the wait includes their lateness)
dfx_status wait_halos(dfx_ssr_strips* s, cudaStream_t st, int slot)
{
    if (!has_neighbours(s)) return DFX_OK;
    DFX_PROFILE(st, "strips_halo_wait");
    flag_wait_kernel<<<1, 1, 0, st>>>(s->up >= 0 ? &s->sync->from_up[slot] : nullptr, s->down >= 0 ? &s->sync->from_down[slot] : nullptr, flag_value(s, slot), &s->sync->error);
    DFX_LAUNCHED("flag_wait_kernel");
    return DFX_OK;
}
dfx_status push_and_wait(dfx_ssr_strips* s, cudaStream_t st, int slot, const PushSpec* specs, int n)
{
    dfx_status rc = push_halos(s, st, slot, specs, n);
    return rc != DFX_OK ? rc : wait_halos(s, st, slot);
}
// A pass whose taps reach `halo` rows beyond the strip, issued so that the neighbours' lateness hides behind useful work: the interior
// rows (which read nothing of the neighbours') first, then the wait for exchange `slot`, then the `halo` rows next to each neighbour.
template <class F>
dfx_status run_with_halo(dfx_ssr_strips* s, cudaStream_t st, int slot, int halo, F&& pass)
{
    const dfx_rows R = s->rows;
    dfx_status     rc;
    if (!has_neighbours(s) || R.y1 - R.y0 <= 2 * halo) // nothing to hide behind
    {
        if ((rc = wait_halos(s, st, slot)) != DFX_OK) return rc;
        return pass(R);
    }
    const int a = s->up >= 0 ? R.y0 + halo : R.y0, b = s->down >= 0 ? R.y1 - halo : R.y1;
    if ((rc = pass(dfx_rows{a, b})) != DFX_OK) return rc;
    if ((rc = wait_halos(s, st, slot)) != DFX_OK) return rc;
    if (a > R.y0 && (rc = pass(dfx_rows{R.y0, a})) != DFX_OK) return rc;
    if (b < R.y1 && (rc = pass(dfx_rows{b, R.y1})) != DFX_OK) return rc;
    return DFX_OK;
}
// The whole Hi-Z pyramid (depth = level 0 included) of this rank's strip, copied into EVERY peer's slab (an all-gather by peer stores;
// the all-rank barrier that follows is its completion). The march descends and climbs the pyramid in a dependent chain of ~36 loads
// per ray, and reflection rays run mostly vertically - out of a row strip: served from the owner over NVLink every one of those loads
// pays the link's latency (measured on 2 GPUs: 1.006x with all levels remote, 1.41x with only level 0 remote - the bottom strip's march
// alone then took longer than the unsharded frame's). 5.33 B/px replicated once per frame; the colour and the normal at the hit (two
// independent loads per ray, after the loop) stay peer loads.
dfx_status gather_hiz_levels(dfx_ssr_strips* s, cudaStream_t st, int first_level, int levels)
{
    if (s->world == 1 || s->rows.y1 <= s->rows.y0) return DFX_OK;
    DFX_PROFILE(st, first_level == 0 ? "strips_gather_depth" : "strips_gather_hiz");
    PushArgs a{};
    for (int m = first_level; m < levels; ++m)
    {
        const dfx_plane& p  = m == 0 ? s->plane[DFX_SSR_STRIPS_PLANE_DEPTH] : s->plane[DFX_SSR_STRIPS_PLANE_HIZ1 + m - 1];
        const int        r0 = s->rows.y0 >> m, r1 = s->rows.y1 >= s->h ? p.height : (s->rows.y1 >> m);
        if (r1 <= r0) continue;
        const int   chunks = int(align_up(size_t(p.width) * 4, 16) / 16);
        const char* b      = static_cast<const char*>(p.ptr);
        for (int r = 0; r < s->world; ++r)
            if (r != s->rank && a.nseg < kMaxPushSegs) a.seg[a.nseg++] = PushSeg{b + size_t(r0) * p.pitch_bytes, s->remote(const_cast<char*>(b), r) - b, (long long)p.pitch_bytes, r1 - r0, chunks};
    }
    a.tickets = &s->sync->push_tickets[kFlagSlots - 1 - (first_level == 0 ? 1 : 0)]; // the two gathers may be in flight together
    a.value   = 0; // no flags: the all-rank barrier orders it
    halo_push_kernel<<<148, 256, 0, st>>>(a);
    DFX_LAUNCHED("halo_push_kernel (Hi-Z all-gather)");
    return DFX_OK;
}

dfx_status all_rank_barrier(dfx_ssr_strips* s, cudaStream_t st, int which)
{
    if (s->world == 1) return DFX_OK;
    DFX_PROFILE(st, which == 0 ? "strips_barrier_before_march" : "strips_barrier_frame_end");
    BarrierArgs a{};
    for (int r = 0; r < s->world; ++r) a.remote[r] = &s->remote(s->sync, r)->barrier[s->rank];
    a.local = s->sync->barrier, a.error = &s->sync->error, a.count = s->world, a.me = s->rank, a.value = s->seq * 2 + unsigned(which);
    barrier_kernel<<<1, 32, 0, st>>>(a);
    DFX_LAUNCHED("barrier_kernel");
    return DFX_OK;
}
} // namespace

// One frame of S1-S7 (+ the PostFX planes they need) on this rank's strip. The caller has written its OWN rows of the six input
// planes (dfx_ssr_strips_plane) in stream order before this call; the SSR output is complete on the own rows when it returns (in
// stream order), and every rank has finished the frame (the call ends with an all-rank barrier), so the inputs may be overwritten.
extern "C" dfx_status dfx_ssr_strips_execute(dfx_ssr_strips* s, void* stream, uint32_t frame_index, const dfx_camera_attribs* curr_camera,
                                             const dfx_camera_attribs* prev_camera, const dfx_ssr_attribs* attribs)
{
    DFX_REQUIRE(s && curr_camera && prev_camera && attribs, "null argument");
    cudaStream_t st = as_stream(stream);
    s->seq += 1;
    const dfx_rows R = s->rows;
    const int      H = s->h;
    auto           P = [&](int id) { return &s->plane[id]; };
    const uint32_t cur = frame_index & 1u, prv = (frame_index + 1u) & 1u;
    dfx_status     rc;

    if ((rc = launch_upload_cameras(st, curr_camera, prev_camera, frame_index, s->cams_dev, nullptr)) != DFX_OK) return rc;

    // E0: the inputs' halos. Depth: 64 rows from below (the Hi-Z rows of the strip's last block reach into the next block wherever a
    // level has an odd height: SSR_ComputeHierarchicalDepthBuffer.fx:52-70 reads row 2y+2), 4 rows from above and 4 more uses below
    // (S5 / S7 taps, S7's quad partner, the 3x3 closest-depth search). Normal, material: +-4 (S5 / S7 taps; S2 is evaluated on the
    // halo rows too instead of exchanging its outputs). Motion: +-1 (closest motion).
    // The depth plane (level 0 of the Hi-Z pyramid, 3/4 of its bytes) is an INPUT: its all-gather starts now, on a second stream, and
    // runs beside the exchange of the input halos and the pre-march passes; the march waits for it through the barrier below.
    if (s->world > 1)
    {
        DFX_CUDA(cudaEventRecord(s->ev_begin, st));
        DFX_CUDA(cudaStreamWaitEvent(s->side, s->ev_begin, 0));
        {
            // one contiguous block of rows per peer, moved by the copy engines (no SM is spent on the 3/4 of the pyramid's bytes)
            DFX_PROFILE(s->side, "strips_gather_depth");
            const dfx_plane& d = s->plane[DFX_SSR_STRIPS_PLANE_DEPTH];
            char*            b = static_cast<char*>(d.ptr) + size_t(R.y0) * d.pitch_bytes;
            const size_t     n = size_t(R.y1 - R.y0) * d.pitch_bytes;
            for (int k = 1; k < s->world && n > 0; ++k)
            {
                const int r = (s->rank + k) % s->world; // every rank starts with a different peer: no link is hit by all at once
                DFX_CUDA(cudaMemcpyAsync(s->remote(b, r), b, n, cudaMemcpyDefault, s->side));
            }
        }
        DFX_CUDA(cudaEventRecord(s->ev_side, s->side));
    }
    const PushSpec e0[] = {{DFX_SSR_STRIPS_PLANE_DEPTH, 64, 4}, {DFX_SSR_STRIPS_PLANE_NORMAL, 4, 4}, {DFX_SSR_STRIPS_PLANE_MATERIAL, 4, 4}, {DFX_SSR_STRIPS_PLANE_MOTION, 1, 1}};
    if ((rc = push_and_wait(s, st, 0, e0, 4)) != DFX_OK) return rc;

    if ((rc = dfx_pass_blue_noise(st, s->tables_dev, frame_index, &s->bn_xy, &s->bn_zw)) != DFX_OK) return rc;
    if ((rc = dfx_pass_postfx_prepare(st, s->cams_dev, P(DFX_SSR_STRIPS_PLANE_DEPTH), P(DFX_SSR_STRIPS_PLANE_PREV_DEPTH_IN), P(DFX_SSR_STRIPS_PLANE_MOTION),
                                      P(DFX_SSR_STRIPS_PLANE_REPROJECTED_DEPTH), P(DFX_SSR_STRIPS_PLANE_CLOSEST_MOTION), P(DFX_SSR_STRIPS_PLANE_PREVIOUS_DEPTH), R)) != DFX_OK)
        return rc;
    dfx_pyramid hz{};
    hz.levels   = 7;
    hz.level[0] = *P(DFX_SSR_STRIPS_PLANE_DEPTH);
    for (int k = 1; k <= 6; ++k) hz.level[k] = *P(DFX_SSR_STRIPS_PLANE_HIZ1 + k - 1);
    {
        int levels = 1; // Hi-Z has min(mip count, 7) levels (ScreenSpaceReflection.cpp:99-133)
        for (int m = std::max(s->w, s->h); m > 1 && levels < 7; m >>= 1) ++levels;
        hz.levels = levels;
    }
    const dfx_rows hiz_rows{R.y0, (s->down >= 0) ? std::min(R.y1 + 64, H) : R.y1};
    if (R.y1 > R.y0 && (rc = dfx_pass_ssr_hiz(st, &hz, hiz_rows)) != DFX_OK) return rc;
    const dfx_rows wide{std::max(R.y0 - 4, 0), std::min(R.y1 + 4, H)};
    if (R.y1 > R.y0 && (rc = dfx_pass_ssr_mask_roughness(st, attribs, P(DFX_SSR_STRIPS_PLANE_MATERIAL), P(DFX_SSR_STRIPS_PLANE_DEPTH), P(DFX_SSR_STRIPS_PLANE_ROUGHNESS),
                                                         P(DFX_SSR_STRIPS_PLANE_MASK), wide)) != DFX_OK)
        return rc;

    // S4 marches on a complete local copy of the Hi-Z pyramid and loads the colour / normal at the hit from whichever rank owns the
    // row, S6 last frame's history: after the barrier everybody's are complete
    if ((rc = gather_hiz_levels(s, st, 1, hz.levels)) != DFX_OK) return rc;
    if (s->world > 1) DFX_CUDA(cudaStreamWaitEvent(st, s->ev_side, 0));
    if ((rc = all_rank_barrier(s, st, 0)) != DFX_OK) return rc;
    if (s->world > 1)
        rc = dfx_pass_ssr_intersect_peer(st, s->cams_dev, attribs, 0, &s->peer_set, P(DFX_SSR_STRIPS_PLANE_COLOR), P(DFX_SSR_STRIPS_PLANE_NORMAL), P(DFX_SSR_STRIPS_PLANE_ROUGHNESS),
                                         P(DFX_SSR_STRIPS_PLANE_MASK), &s->bn_xy, &hz, P(DFX_SSR_STRIPS_PLANE_RADIANCE), P(DFX_SSR_STRIPS_PLANE_RAYDIR), R);
    else
        rc = dfx_pass_ssr_intersect(st, s->cams_dev, attribs, 0, P(DFX_SSR_STRIPS_PLANE_COLOR), P(DFX_SSR_STRIPS_PLANE_NORMAL), P(DFX_SSR_STRIPS_PLANE_ROUGHNESS),
                                    P(DFX_SSR_STRIPS_PLANE_MASK), &s->bn_xy, &hz, P(DFX_SSR_STRIPS_PLANE_MOTION), P(DFX_SSR_STRIPS_PLANE_RADIANCE), P(DFX_SSR_STRIPS_PLANE_RAYDIR), R);
    if (rc != DFX_OK) return rc;

    // S5: 8-tap disk of radius <= 4 px over the ray planes
    const PushSpec e1[] = {{DFX_SSR_STRIPS_PLANE_RADIANCE, 4, 4}, {DFX_SSR_STRIPS_PLANE_RAYDIR, 4, 4}};
    if ((rc = push_halos(s, st, 1, e1, 2)) != DFX_OK) return rc;
    if ((rc = run_with_halo(s, st, 1, 4, [&](dfx_rows rows) {
             return dfx_pass_ssr_spatial(st, s->cams_dev, attribs, P(DFX_SSR_STRIPS_PLANE_ROUGHNESS), P(DFX_SSR_STRIPS_PLANE_MASK), P(DFX_SSR_STRIPS_PLANE_NORMAL), P(DFX_SSR_STRIPS_PLANE_DEPTH),
                                         P(DFX_SSR_STRIPS_PLANE_RAYDIR), P(DFX_SSR_STRIPS_PLANE_RADIANCE), P(DFX_SSR_STRIPS_PLANE_RESOLVED_RADIANCE),
                                         P(DFX_SSR_STRIPS_PLANE_RESOLVED_VARIANCE), P(DFX_SSR_STRIPS_PLANE_RESOLVED_DEPTH), rows);
         })) != DFX_OK)
        return rc;

    // S6: 3x3 statistics of the resolved radiance (halo +-1); last frame's planes through peer loads
    const PushSpec e2[] = {{DFX_SSR_STRIPS_PLANE_RESOLVED_RADIANCE, 1, 1}};
    if ((rc = push_halos(s, st, 2, e2, 1)) != DFX_OK) return rc;
    const int rh_prv = prv ? DFX_SSR_STRIPS_PLANE_RADIANCE_HISTORY1 : DFX_SSR_STRIPS_PLANE_RADIANCE_HISTORY0, rh_cur = cur ? DFX_SSR_STRIPS_PLANE_RADIANCE_HISTORY1 : DFX_SSR_STRIPS_PLANE_RADIANCE_HISTORY0;
    const int vh_prv = prv ? DFX_SSR_STRIPS_PLANE_VARIANCE_HISTORY1 : DFX_SSR_STRIPS_PLANE_VARIANCE_HISTORY0, vh_cur = cur ? DFX_SSR_STRIPS_PLANE_VARIANCE_HISTORY1 : DFX_SSR_STRIPS_PLANE_VARIANCE_HISTORY0;
    if ((rc = run_with_halo(s, st, 2, 1, [&](dfx_rows rows) {
             if (s->world > 1)
                 return dfx_pass_ssr_temporal_peer(st, s->cams_dev, attribs, &s->peers, P(DFX_SSR_STRIPS_PLANE_MASK), P(DFX_SSR_STRIPS_PLANE_MOTION), P(DFX_SSR_STRIPS_PLANE_RESOLVED_DEPTH),
                                                   P(DFX_SSR_STRIPS_PLANE_REPROJECTED_DEPTH), P(DFX_SSR_STRIPS_PLANE_RESOLVED_RADIANCE), P(DFX_SSR_STRIPS_PLANE_RESOLVED_VARIANCE),
                                                   P(DFX_SSR_STRIPS_PLANE_PREVIOUS_DEPTH), P(rh_prv), P(vh_prv), P(rh_cur), P(vh_cur), rows);
             return dfx_pass_ssr_temporal(st, s->cams_dev, attribs, P(DFX_SSR_STRIPS_PLANE_MASK), P(DFX_SSR_STRIPS_PLANE_MOTION), P(DFX_SSR_STRIPS_PLANE_RESOLVED_DEPTH),
                                          P(DFX_SSR_STRIPS_PLANE_REPROJECTED_DEPTH), P(DFX_SSR_STRIPS_PLANE_RESOLVED_RADIANCE), P(DFX_SSR_STRIPS_PLANE_RESOLVED_VARIANCE),
                                          P(DFX_SSR_STRIPS_PLANE_PREVIOUS_DEPTH), P(rh_prv), P(vh_prv), P(rh_cur), P(vh_cur), rows);
         })) != DFX_OK)
        return rc;

    // S7: (2r+1)^2 window, r <= 2, over this frame's radiance history
    const PushSpec e3[] = {{rh_cur, 2, 2}};
    if ((rc = push_halos(s, st, 3, e3, 1)) != DFX_OK) return rc;
    if ((rc = run_with_halo(s, st, 3, 2, [&](dfx_rows rows) {
             return dfx_pass_ssr_bilateral(st, s->cams_dev, attribs, P(DFX_SSR_STRIPS_PLANE_MASK), P(DFX_SSR_STRIPS_PLANE_DEPTH), P(DFX_SSR_STRIPS_PLANE_NORMAL), P(DFX_SSR_STRIPS_PLANE_ROUGHNESS),
                                           P(rh_cur), P(vh_cur), P(DFX_SSR_STRIPS_PLANE_OUTPUT), rows);
         })) != DFX_OK)
        return rc;
    // nobody overwrites an input, a Hi-Z level or a history slot that a peer may still be loading from
    return all_rank_barrier(s, st, 1);
}
