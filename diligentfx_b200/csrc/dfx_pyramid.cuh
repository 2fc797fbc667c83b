// dfx_pyramid.cuh — 2:1 pyramids (SSR Hi-Z S1, SSAO prefiltered depth A2, SSAO convoluted AO + depth A6) built in one or two
// launches instead of the reference's one draw per level (ScreenSpaceReflection.cpp:777-902, ScreenSpaceAmbientOcclusion.cpp:
// 881-957, :1075-1255).
//
// The per-texel rule is the shaders' (cited at each Op): level m texel (x, y) reduces the level m-1 texels (2x..2x+1, 2y..2y+1),
// plus column 2x+2 when the source width is odd and row 2y+2 when the source height is odd, in the shaders' ArrayAppend order.
//
//   * tile kernel  — while the source dimensions stay even (levels 1..K, K = trailing zero bits of the frame size), a level-m
//     texel depends on exactly one 2^m x 2^m block of level 0. A CTA takes a 64x64 block of level 0, staged into shared memory
//     by ONE TMA 2-D tile load per plane (cp.async.bulk.tensor, zero fill beyond the plane), and writes every level up to K from
//     registers / warp shuffles / shared memory: level 0 is read from HBM exactly once, no intermediate level is re-read;
//   * tail kernel  — the remaining levels (odd-sized sources, <= 16K texels each at 4K / 8K) in one thread-block-cluster launch,
//     a cluster barrier between levels;
//   * level kernel — the generic one-launch-per-level form, for row strips (multi-GPU) and for sizes the two above do not cover.
// All three evaluate the same Op::reduce on the same values: results are bit-identical whichever path builds a level.
#pragma once
#include "dfx_common.cuh"
#include "dfx_tma.cuh"
#include <cooperative_groups.h>

namespace dfx
{

constexpr int kPyrTile        = 64;
constexpr int kPyrTailTexels  = 16384;
constexpr int kPyrTailThreads = 512;
constexpr int kPyrTailCluster = 8;

template <int N>
struct PyrVal
{
    float v[N];
};

template <int N>
struct PyrPlanes
{
    View<float> lv[N][DFX_MAX_MIPS]; // [plane][level]; level 0 is only read
    int         levels;
};

template <int N>
struct PyrMaps
{
    CUtensorMap m[N];
};

// ---- the three reductions -------------------------------------------------------------------------------------------------------
// S1, SSR_ComputeHierarchicalDepthBuffer.fx:30-73: closest depth (min; max with reversed depth), seeded with the far plane
struct HizOp
{
    static constexpr int N = 1;
    int                  rev;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ PyrVal<1> reduce(const PyrVal<1>* t, int n) const
    {
        float m = rev ? 0.0f : 1.0f;
        for (int i = 0; i < n; ++i) m = rev ? fmaxf(m, t[i].v[0]) : fminf(m, t[i].v[0]);
        return PyrVal<1>{{m}};
    }
};

// A2, SSAO_ComputePrefilteredDepthBuffer.fx:42-71, :79-122: taps to view-space Z, weighted average favouring the closest tap within
// the falloff range, back to depth, saturate
struct PrefilterOp
{
    static constexpr int      N = 1;
    const dfx_camera_attribs* cams;
    float                     falloffMul, falloffAdd;
    CamS                      cam;
    __device__ __forceinline__ void init() { load_cam(cam, &cams[0]); }
    __device__ __forceinline__ PyrVal<1> reduce(const PyrVal<1>* t, int n) const
    {
        float z[9];
        for (int i = 0; i < n; ++i) z[i] = depth_to_camz(t[i].v[0], cam);
        float zmin = z[0];
        for (int i = 1; i < n; ++i) zmin = fminf(zmin, z[i]);
        float zsum = 0.0f, wsum = 0.0f;
        for (int i = 0; i < n; ++i)
        {
            const float w = saturate(fabsf(zmin - z[i]) * falloffMul + falloffAdd);
            zsum += w * z[i];
            wsum += w;
        }
        return PyrVal<1>{{saturate(camz_to_depth(zsum / wsum, cam))}};
    }
};
inline PrefilterOp make_prefilter_op(const dfx_camera_attribs* cams_dev, const dfx_ssao_attribs& A)
{
    PrefilterOp op;
    op.cams                  = cams_dev;
    const float radius       = 0.75f * A.EffectRadius * A.RadiusMultiplier;
    const float falloffRange = A.EffectFalloffRange * radius;
    const float falloffFrom  = radius - falloffRange;
    op.falloffMul            = -1.0f / falloffRange;
    op.falloffAdd            = falloffFrom / falloffRange + 1.0f;
    return op;
}

// A6, SSAO_ComputeConvolutedDepthHistory.fx:93-109: plain averages of the AO history and of the depth, accumulated in tap order
struct ConvoluteOp
{
    static constexpr int N = 2;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ PyrVal<2> reduce(const PyrVal<2>* t, int n) const
    {
        float a = t[0].v[0], b = t[0].v[1];
        for (int i = 1; i < n; ++i) a += t[i].v[0], b += t[i].v[1];
        return PyrVal<2>{{a / float(n), b / float(n)}};
    }
};

// taps of level-m texel (x, y) in level m-1, in ArrayAppend order: (0,0) (0,1) (1,0) (1,1) [w odd: (2,0) (2,1)] [h odd: (0,2) (1,2)] [both: (2,2)]
template <class Op, bool CG>
__device__ __forceinline__ PyrVal<Op::N> pyr_reduce_texel(const Op& op, const PyrPlanes<Op::N>& P, int m, int x, int y)
{
    constexpr int N = Op::N;
    const int     sw = P.lv[0][m - 1].w, sh = P.lv[0][m - 1].h;
    const bool    wodd = sw & 1, hodd = sh & 1;
    PyrVal<N>     t[9];
    int           n = 0;
    auto tap = [&](int ox, int oy) {
        const int tx = min(2 * x + ox, sw - 1), ty = min(2 * y + oy, sh - 1);
#pragma unroll
        for (int k = 0; k < N; ++k)
        {
            const float* p = &P.lv[k][m - 1].at(tx, ty);
            t[n].v[k]      = CG ? __ldcg(p) : __ldg(p);
        }
        ++n;
    };
    tap(0, 0), tap(0, 1), tap(1, 0), tap(1, 1);
    if (wodd) tap(2, 0), tap(2, 1);
    if (hodd) tap(0, 2), tap(1, 2);
    if (wodd && hodd) tap(2, 2);
    return op.reduce(t, n);
}

// ---- generic: one launch per level (rows r0..r1 of level m) ----------------------------------------------------------------------
template <class Op>
__global__ void __launch_bounds__(256) pyramid_level_kernel(Op op, PyrPlanes<Op::N> P, int m, int r0, int r1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = r0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.lv[0][m].w || y >= r1) return;
    op.init();
    const PyrVal<Op::N> r = pyr_reduce_texel<Op, false>(op, P, m, x, y);
#pragma unroll
    for (int k = 0; k < Op::N; ++k) P.lv[k][m].at(x, y) = r.v[k];
}

// ---- tail: levels m0..m1 in one cluster launch ------------------------------------------------------------------------------------
template <class Op>
__global__ void __cluster_dims__(kPyrTailCluster, 1, 1) __launch_bounds__(kPyrTailThreads) pyramid_tail_kernel(Op op, PyrPlanes<Op::N> P, int m0, int m1)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int         tid = int(cluster.block_rank()) * kPyrTailThreads + int(threadIdx.x), nth = kPyrTailCluster * kPyrTailThreads;
    op.init();
    for (int m = m0; m <= m1; ++m)
    {
        const int w = P.lv[0][m].w, h = P.lv[0][m].h;
        for (int idx = tid; idx < w * h; idx += nth)
        {
            const int           y = idx / w, x = idx - y * w;
            const PyrVal<Op::N> r = pyr_reduce_texel<Op, true>(op, P, m, x, y); // ld.global.cg: the source may have been written by another CTA of this launch
#pragma unroll
            for (int k = 0; k < Op::N; ++k) P.lv[k][m].at(x, y) = r.v[k];
        }
        if (m < m1) cluster.sync();
    }
}

// ---- tile: levels 1..K from one TMA-staged 64x64 block of level 0 per CTA -------------------------------------------------------
// 256 threads as 16x16; thread (tx, ty) owns the 4x4 level-0 block at (4tx, 4ty) of the tile: levels 1 and 2 in registers, level 3
// by warp shuffles (a warp holds two thread rows: the 2x2 partners are lane ^ 1 and lane ^ 16), levels 4..6 through shared memory.
template <class Op, bool USE_TMA>
__global__ void __launch_bounds__(256) pyramid_tile_kernel(Op op, const __grid_constant__ PyrMaps<Op::N> maps, PyrPlanes<Op::N> P, int K, int y0)
{
    constexpr int N = Op::N;
    __shared__ __align__(128) float tile[N][kPyrTile][kPyrTile];
    __shared__ float                s3[N][8][8], s4[N][4][4], s5[N][2][2];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int X0 = blockIdx.x * kPyrTile, Y0 = y0 + blockIdx.y * kPyrTile;
    if (USE_TMA)
    {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0)
        {
            mbar_arrive_expect_tx(&bar, uint32_t(sizeof(tile)));
#pragma unroll
            for (int k = 0; k < N; ++k) tma_load_2d(&tile[k][0][0], &maps.m[k], X0, Y0, &bar);
        }
        op.init(); // camera constants arrive while the tile is in flight
        mbar_wait(&bar, 0);
    }
    else
    {
        op.init();
        const int W = P.lv[0][0].w, H = P.lv[0][0].h;
#pragma unroll
        for (int k = 0; k < N; ++k)
            for (int i = tid; i < kPyrTile * kPyrTile; i += 256)
            {
                const int ly = i >> 6, lx = i & 63;
                tile[k][ly][lx] = (X0 + lx < W && Y0 + ly < H) ? __ldg(&P.lv[k][0].at(X0 + lx, Y0 + ly)) : 0.0f;
            }
        __syncthreads();
    }

    // level 0 block -> registers (4 x LDS.128 per plane; a quarter warp reads 128 contiguous bytes: conflict-free)
    float b[N][4][4];
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            const float4 q = *reinterpret_cast<const float4*>(&tile[k][4 * ty + r][4 * tx]);
            b[k][r][0] = q.x, b[k][r][1] = q.y, b[k][r][2] = q.z, b[k][r][3] = q.w;
        }
    auto reduce4 = [&](const PyrVal<N>& v00, const PyrVal<N>& v01, const PyrVal<N>& v10, const PyrVal<N>& v11) {
        const PyrVal<N> t[4] = {v00, v01, v10, v11}; // (ox, oy) = (0,0) (0,1) (1,0) (1,1)
        return op.reduce(t, 4);
    };
    auto store = [&](int m, int x, int y, const PyrVal<N>& v) {
        if (x < P.lv[0][m].w && y < P.lv[0][m].h)
#pragma unroll
            for (int k = 0; k < N; ++k) P.lv[k][m].at(x, y) = v.v[k];
    };
    auto at0 = [&](int r, int c) {
        PyrVal<N> v;
#pragma unroll
        for (int k = 0; k < N; ++k) v.v[k] = b[k][r][c];
        return v;
    };
    // level 1: 2x2 per thread
    PyrVal<N> l1[2][2]; // [j = y][i = x]
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i)
        {
            l1[j][i] = reduce4(at0(2 * j, 2 * i), at0(2 * j + 1, 2 * i), at0(2 * j, 2 * i + 1), at0(2 * j + 1, 2 * i + 1));
            store(1, (X0 >> 1) + 2 * tx + i, (Y0 >> 1) + 2 * ty + j, l1[j][i]);
        }
    if (K < 2) return;
    // level 2: one per thread
    const PyrVal<N> l2 = reduce4(l1[0][0], l1[1][0], l1[0][1], l1[1][1]);
    store(2, (X0 >> 2) + tx, (Y0 >> 2) + ty, l2);
    if (K < 3) return;
    // level 3: the 2x2 partners sit in this warp (lane = (ty & 1) * 16 + tx)
    const int lane = tid & 31, base = lane & ~17;
    PyrVal<N> q00, q01, q10, q11;
#pragma unroll
    for (int k = 0; k < N; ++k)
    {
        q00.v[k] = __shfl_sync(0xffffffffu, l2.v[k], base);
        q01.v[k] = __shfl_sync(0xffffffffu, l2.v[k], base | 16);
        q10.v[k] = __shfl_sync(0xffffffffu, l2.v[k], base | 1);
        q11.v[k] = __shfl_sync(0xffffffffu, l2.v[k], base | 17);
    }
    const PyrVal<N> l3 = reduce4(q00, q01, q10, q11);
    if (lane == base)
    {
        store(3, (X0 >> 3) + (tx >> 1), (Y0 >> 3) + (ty >> 1), l3);
#pragma unroll
        for (int k = 0; k < N; ++k) s3[k][ty >> 1][tx >> 1] = l3.v[k];
    }
    if (K < 4) return;
    __syncthreads();
    if (tid < 16) // level 4: 4x4 per tile
    {
        const int x = tid & 3, y = tid >> 2;
        PyrVal<N> t[4];
#pragma unroll
        for (int k = 0; k < N; ++k) t[0].v[k] = s3[k][2 * y][2 * x], t[1].v[k] = s3[k][2 * y + 1][2 * x], t[2].v[k] = s3[k][2 * y][2 * x + 1], t[3].v[k] = s3[k][2 * y + 1][2 * x + 1];
        const PyrVal<N> r = op.reduce(t, 4);
        store(4, (X0 >> 4) + x, (Y0 >> 4) + y, r);
#pragma unroll
        for (int k = 0; k < N; ++k) s4[k][y][x] = r.v[k];
    }
    if (K < 5) return;
    __syncthreads();
    if (tid < 4) // level 5: 2x2 per tile
    {
        const int x = tid & 1, y = tid >> 1;
        PyrVal<N> t[4];
#pragma unroll
        for (int k = 0; k < N; ++k) t[0].v[k] = s4[k][2 * y][2 * x], t[1].v[k] = s4[k][2 * y + 1][2 * x], t[2].v[k] = s4[k][2 * y][2 * x + 1], t[3].v[k] = s4[k][2 * y + 1][2 * x + 1];
        const PyrVal<N> r = op.reduce(t, 4);
        store(5, (X0 >> 5) + x, (Y0 >> 5) + y, r);
#pragma unroll
        for (int k = 0; k < N; ++k) s5[k][y][x] = r.v[k];
    }
    if (K < 6) return;
    __syncthreads();
    if (tid == 0) // level 6: one per tile
    {
        PyrVal<N> t[4];
#pragma unroll
        for (int k = 0; k < N; ++k) t[0].v[k] = s5[k][0][0], t[1].v[k] = s5[k][1][0], t[2].v[k] = s5[k][0][1], t[3].v[k] = s5[k][1][1];
        store(6, X0 >> 6, Y0 >> 6, op.reduce(t, 4));
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------------
const CUtensorMap* tensor_map_r32f(const View<float>& plane, int box_w, int box_h); // cached per (pointer, size, pitch, box); nullptr if the plane cannot be described
int                tune(const char* name, int fallback);
bool               async_compute_hint(); // the chain executor is issuing a frame whose SSR and SSAO halves share the GPU (dfx_api.cu)

inline int trailing_zeros(int v)
{
    int n = 0;
    while (v > 0 && (v & 1) == 0) ++n, v >>= 1;
    return n;
}
inline int pyr_mip_row(int y, int m, int full_h, int mip_h) { return y >= full_h ? mip_h : (y >> m); }

// Builds levels 1..max_m (clamped to the pyramid) over the level-0 rows [rows.y0, rows.y1). `launches` returns how many kernels ran.
template <class Op>
dfx_status build_pyramid(void* stream, const Op& op, const PyrPlanes<Op::N>& P, int max_m, dfx_rows rows, const char* what)
{
    constexpr int N = Op::N;
    cudaStream_t  s = as_stream(stream);
    const int     W = P.lv[0][0].w, H = P.lv[0][0].h;
    max_m           = min(max_m, P.levels - 1);
    const bool full = rows.y0 == 0 && rows.y1 == H;
    int        m    = 1;
    // 2 = TMA tile + cluster tail, 1 = tile staged with plain loads + tail, 0 = one launch per level. As an isolated pass the tile kernel is
    // twice as fast as the per-level launches (0.019 vs 0.040 ms for the 4K Hi-Z), but with the SSR and SSAO halves of the frame side by
    // side on two streams the FRAME is 1.7 % faster with per-level launches (2.218 vs 2.255 ms, profiles/r2k1b): a tile CTA holds 256
    // threads and 16-32 KB of shared memory while ever fewer of its threads reduce the upper levels, and that residency is taken from
    // the issue-bound kernel of the other stream; small CTAs that come and go share the SMs better. So the default follows the issuer.
    const int  mode = tune("pyramid_impl", async_compute_hint() ? 0 : 2);
    if (full && mode != 0 && max_m >= 1)
    {
        const int K = min(min(trailing_zeros(W), trailing_zeros(H)), min(max_m, 6));
        if (K >= 2)
        {
            const dim3 grid(div_up(W, kPyrTile), div_up(H, kPyrTile));
            PyrMaps<N> maps{};
            bool       tma = mode == 2;
            for (int k = 0; k < N && tma; ++k)
            {
                const CUtensorMap* tm = tensor_map_r32f(P.lv[k][0], kPyrTile, kPyrTile);
                if (tm)
                    maps.m[k] = *tm;
                else
                    tma = false;
            }
            if (tma)
                pyramid_tile_kernel<Op, true><<<grid, 256, 0, s>>>(op, maps, P, K, 0);
            else
                pyramid_tile_kernel<Op, false><<<grid, 256, 0, s>>>(op, maps, P, K, 0);
            DFX_LAUNCHED(what);
            m = K + 1;
        }
    }
    for (; m <= max_m; ++m)
    {
        const int lw = P.lv[0][m].w, lh = P.lv[0][m].h;
        if (full && mode != 0 && tune("pyramid_tail", 1) && (long long)lw * lh <= kPyrTailTexels)
        {
            pyramid_tail_kernel<Op><<<kPyrTailCluster, kPyrTailThreads, 0, s>>>(op, P, m, max_m);
            DFX_LAUNCHED(what);
            break;
        }
        const int r0 = pyr_mip_row(rows.y0, m, H, lh), r1 = pyr_mip_row(rows.y1, m, H, lh);
        if (r1 <= r0) continue;
        const dim3 block(32, 8), grid(div_up(lw, 32), div_up(r1 - r0, 8));
        pyramid_level_kernel<Op><<<grid, block, 0, s>>>(op, P, m, r0, r1);
        DFX_LAUNCHED(what);
    }
    return DFX_OK;
}

} // namespace dfx
