// dfx_bloom_taa_tonemap.cu — Bloom B1-B4, TAA T1, compose, ToneMap M1/M2 as sm_100a kernels.
// Reference: PostProcess/Bloom/src/Bloom.cpp:288-393 + Shaders/PostProcess/Bloom/private/*.fx;
//            PostProcess/TemporalAntiAliasing/src/TemporalAntiAliasing.cpp:260-289 + TAA_ComputeTemporalAccumulation.fx;
//            Shaders/PostProcess/ToneMapping/public/ToneMapping.fxh; Hydrogent/shaders/HnPostProcess.psh:145-185, HnCopyFrame.psh:32-62.
#include "dfx_common.cuh"
#ifndef DFX_BLOOM_TMA
#    define DFX_BLOOM_TMA 0
#endif
#if DFX_BLOOM_TMA
#    include "dfx_tma.cuh"
#endif

namespace dfx
{

// =====================================================================================================================
// Bloom. Levels are RGBA32F planes (rgb used). B1/B2 sample with linear + border(0) addressing, B3/B4 with linear + clamp.
// The output pixel centre in UV is (p + 0.5) / output size; taps are offset by whole input texels.
// =====================================================================================================================
template <bool BORDER>
DFX_HD float3 tap3(const View<const float4>& t, float u, float v)
{
    return xyz(BORDER ? sample_linear_border(t, u, v) : sample_linear_clamp(t, u, v));
}

struct Taps13
{
    float3 A, B, C, D, E, F, G, H, I, J, K, L, M;
};
DFX_HD Taps13 taps13(const View<const float4>& in, float u, float v)
{
    const float tx = 1.0f / float(in.w), ty = 1.0f / float(in.h);
    Taps13      t;
    t.A = tap3<true>(in, u + tx * -2.0f, v + ty * +2.0f);
    t.B = tap3<true>(in, u + tx * +0.0f, v + ty * +2.0f);
    t.C = tap3<true>(in, u + tx * +2.0f, v + ty * +2.0f);
    t.D = tap3<true>(in, u + tx * -2.0f, v + ty * +0.0f);
    t.E = tap3<true>(in, u + tx * +0.0f, v + ty * +0.0f);
    t.F = tap3<true>(in, u + tx * +2.0f, v + ty * +0.0f);
    t.G = tap3<true>(in, u + tx * -2.0f, v + ty * -2.0f);
    t.H = tap3<true>(in, u + tx * +0.0f, v + ty * -2.0f);
    t.I = tap3<true>(in, u + tx * +2.0f, v + ty * -2.0f);
    t.J = tap3<true>(in, u + tx * -1.0f, v + ty * +1.0f);
    t.K = tap3<true>(in, u + tx * +1.0f, v + ty * +1.0f);
    t.L = tap3<true>(in, u + tx * -1.0f, v + ty * -1.0f);
    t.M = tap3<true>(in, u + tx * +1.0f, v + ty * -1.0f);
    return t;
}

// B1: Bloom_ComputePrefilteredTexture.fx:37-83 — 13 taps in 5 Karis-weighted groups, soft-knee threshold
__global__ void __launch_bounds__(256) bloom_prefilter_kernel(dfx_bloom_attribs A, View<const float4> in, View<float4> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float  u = (float(x) + 0.5f) / float(out.w), v = (float(y) + 0.5f) / float(out.h);
    const Taps13 t = taps13(in, u, v);
    float3       g[5];
    g[0] = (t.A + t.B + t.D + t.E) / 4.0f;
    g[1] = (t.B + t.C + t.E + t.F) / 4.0f;
    g[2] = (t.D + t.E + t.G + t.H) / 4.0f;
    g[3] = (t.E + t.F + t.H + t.I) / 4.0f;
    g[4] = (t.J + t.K + t.L + t.M) / 4.0f;
    float3 csum = make_float3(0.f, 0.f, 0.f);
    float  wsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
    {
        const float w = (i == 4 ? 0.5f : 0.125f) * (1.0f / (1.0f + luminance(g[i])));
        csum = csum + g[i] * w;
        wsum += 1.0f * w;
    }
    const float3 c = csum / (wsum + 1.0e-5f);
    // Prefilter :24-35
    const float brightness = fmaxf(c.x, fmaxf(c.y, c.z));
    const float knee       = A.Threshold * A.SoftTreshold;
    float       soft       = brightness - A.Threshold + knee;
    soft                   = fminf(fmaxf(soft, 0.0f), 2.0f * knee);
    soft                   = soft * soft * 0.25f / (knee + 1.0e-5f);
    float contribution     = fmaxf(soft, brightness - A.Threshold);
    contribution /= fmaxf(brightness, 1.0e-5f);
    out.at(x, y) = f4(c * contribution, 0.0f);
}

// B2: Bloom_ComputeDownsampledTexture.fx:11-41 — 13-tap downsample, weights 1/32, 1/16, 1/8
__global__ void __launch_bounds__(256) bloom_downsample_kernel(View<const float4> in, View<float4> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float  u = (float(x) + 0.5f) / float(out.w), v = (float(y) + 0.5f) / float(out.h);
    const Taps13 t = taps13(in, u, v);
    float3       o = make_float3(0.f, 0.f, 0.f);
    o = o + (t.A + t.C + t.G + t.I) * 0.03125f;
    o = o + (t.B + t.D + t.F + t.H) * 0.0625f;
    o = o + (t.E + t.J + t.K + t.L + t.M) * 0.125f;
    out.at(x, y) = f4(o, 0.0f);
}

DFX_HD float3 tent9(const View<const float4>& lo, float u, float v)
{
    const float  tx = 1.0f / float(lo.w), ty = 1.0f / float(lo.h);
    const float3 A = tap3<false>(lo, u - tx, v + ty), B = tap3<false>(lo, u, v + ty), C = tap3<false>(lo, u + tx, v + ty);
    const float3 D = tap3<false>(lo, u - tx, v), E = tap3<false>(lo, u, v), F = tap3<false>(lo, u + tx, v);
    const float3 G = tap3<false>(lo, u - tx, v - ty), H = tap3<false>(lo, u, v - ty), I = tap3<false>(lo, u + tx, v - ty);
    float3       s = E * 0.25f;
    s = s + (B + D + F + H) * 0.125f;
    s = s + (A + C + G + I) * 0.0625f;
    return s;
}

// B3: Bloom_ComputeUpsampledTexture.fx:20-54 (uInstID == 0): same-level downsample + 3x3 tent of the coarser level
__global__ void __launch_bounds__(256) bloom_upsample_kernel(View<const float4> same, View<const float4> coarser, View<float4> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float  u = (float(x) + 0.5f) / float(out.w), v = (float(y) + 0.5f) / float(out.h);
    const float3 s = tent9(coarser, u, v);
    const float3 c = tap3<false>(same, u, v);
    out.at(x, y)   = f4(c + s, 0.0f);
}

// B4: final composite (uInstID != 0), :45-48
__global__ void __launch_bounds__(256) bloom_composite_kernel(dfx_bloom_attribs A, View<const float4> color, View<const float4> up0,
                                                              View<float4> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float  u = (float(x) + 0.5f) / float(out.w), v = (float(y) + 0.5f) / float(out.h);
    const float3 s = tent9(up0, u, v);
    const float3 c = tap3<false>(color, u, v);
    st_cs(&out.at(x, y), f4(lerp3(c, c + A.Intensity * s, A.AlphaInterpolation), 0.0f));
}

// ---------------------------------------------------------------------------------------------------------------------
// Exact-2:1 fast paths. When the finer plane is exactly twice the coarser one in both dimensions (every large level of the
// pyramid: 3840x2160 -> 1920x1080 -> 960x540 -> 480x270 -> 240x135), all sample positions fall on texel corners
// (down-sampling) or on quarter-texel offsets (up-sampling), exactly representable in the sampler's 8 sub-texel bits. The
// bilinear weights are then the constants 1/4 (corner average) resp. {1/4, 3/4}, so the taps can be evaluated from a
// shared-memory tile of the source with fixed weights instead of 13x4 / 9x4 gathered texels per pixel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDnTileW = 68, kDnTileH = 20; // source texels staged for a 32x8 output tile of B1/B2: (2*32 + 4) x (2*8 + 4)
constexpr int kDnCornW = 67, kDnCornH = 19; // texel corners inside that tile

// B1 / B2 on an exact 2:1 level. Stage 1: the CTA stages the 68x20 source tile (out-of-range texels = 0: border addressing).
// Stage 2: the 67x19 corner averages (each bilinear tap of the shader IS one corner average). Stage 3: 13 taps per output.
// stages 2 and 3, from a staged source tile
template <bool PREFILTER>
__device__ __forceinline__ void bloom_down2x_from_tile(const dfx_bloom_attribs& A, const float4 (&tile)[kDnTileH][kDnTileW], float4 (&corner)[kDnCornH][kDnCornW],
                                                       View<float4> out, int ox0, int oy0, int y1)
{
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < kDnCornW * kDnCornH; i += 256)
    {
        const int    ly = i / kDnCornW, lx = i - ly * kDnCornW;
        const float4 a = tile[ly][lx], b = tile[ly][lx + 1], c = tile[ly + 1][lx], d = tile[ly + 1][lx + 1];
        corner[ly][lx] = (a * 0.25f + b * 0.25f) + (c * 0.25f + d * 0.25f);
    }
    __syncthreads();
    const int x = ox0 + threadIdx.x, y = oy0 + threadIdx.y;
    if (x >= out.w || y >= y1) return;
    // output centre = corner (2*lx + 2, 2*ly + 2) of the tile; tap offset (i, j) texels -> corner (cx + i, cy + j)
    const int cx = 2 * threadIdx.x + 2, cy = 2 * threadIdx.y + 2;
    auto      T  = [&](int i, int j) { return xyz(corner[cy + j][cx + i]); };
    const float3 tA = T(-2, +2), tB = T(0, +2), tC = T(+2, +2), tD = T(-2, 0), tE = T(0, 0), tF = T(+2, 0), tG = T(-2, -2), tH = T(0, -2), tI = T(+2, -2);
    const float3 tJ = T(-1, +1), tK = T(+1, +1), tL = T(-1, -1), tM = T(+1, -1);
    if (!PREFILTER)
    {
        float3 o = make_float3(0.f, 0.f, 0.f);
        o = o + (tA + tC + tG + tI) * 0.03125f;
        o = o + (tB + tD + tF + tH) * 0.0625f;
        o = o + (tE + tJ + tK + tL + tM) * 0.125f;
        out.at(x, y) = f4(o, 0.0f);
        return;
    }
    float3 g[5];
    g[0] = (tA + tB + tD + tE) * 0.25f;
    g[1] = (tB + tC + tE + tF) * 0.25f;
    g[2] = (tD + tE + tG + tH) * 0.25f;
    g[3] = (tE + tF + tH + tI) * 0.25f;
    g[4] = (tJ + tK + tL + tM) * 0.25f;
    float3 csum = make_float3(0.f, 0.f, 0.f);
    float  wsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
    {
        const float w = (i == 4 ? 0.5f : 0.125f) * frcp(1.0f + luminance(g[i]));
        csum = csum + g[i] * w;
        wsum += w;
    }
    const float3 c = csum * frcp(wsum + 1.0e-5f);
    const float brightness = fmaxf(c.x, fmaxf(c.y, c.z));
    const float knee       = A.Threshold * A.SoftTreshold;
    float       soft       = fminf(fmaxf(brightness - A.Threshold + knee, 0.0f), 2.0f * knee);
    soft                   = soft * soft * 0.25f * frcp(knee + 1.0e-5f);
    const float contribution = fmaxf(soft, brightness - A.Threshold) * frcp(fmaxf(brightness, 1.0e-5f));
    out.at(x, y) = f4(c * contribution, 0.0f);
}

template <bool PREFILTER>
__global__ void __launch_bounds__(256) bloom_down2x_kernel(dfx_bloom_attribs A, View<const float4> in, View<float4> out, int y0, int y1)
{
    __shared__ float4 tile[kDnTileH][kDnTileW];
    __shared__ float4 corner[kDnCornH][kDnCornW];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int ox0 = blockIdx.x * 32, oy0 = y0 + blockIdx.y * 8;
    const int sx0 = 2 * ox0 - 2, sy0 = 2 * oy0 - 2;
    for (int i = tid; i < kDnTileW * kDnTileH; i += 256)
    {
        const int ly = i / kDnTileW, lx = i - ly * kDnTileW;
        tile[ly][lx] = load0(in, sx0 + lx, sy0 + ly);
    }
    __syncthreads();
    bloom_down2x_from_tile<PREFILTER>(A, tile, corner, out, ox0, oy0, y1);
}

#if DFX_BLOOM_TMA
// The same pass with the source tile staged by ONE TMA 2-D tile load instead of 1360 predicated 128-bit loads (the north
// star's prescription for the Bloom pyramid). The box starts at texel (2*ox0 - 2, 2*oy0 - 2); the parts of it that lie outside
// the plane arrive as zeros, which is the border(0) addressing of these taps. Everything after the staging is shared with
// bloom_down2x_kernel, so the two produce identical planes. Opt-in build (-DDFX_BLOOM_TMA=1): not yet timed on the GPU.
template <bool PREFILTER>
__global__ void __launch_bounds__(256) bloom_down2x_tma_kernel(dfx_bloom_attribs A, const __grid_constant__ CUtensorMap in_map, View<float4> out, int y0, int y1)
{
    __shared__ __align__(128) float4 tile[kDnTileH][kDnTileW];
    __shared__ float4                corner[kDnCornH][kDnCornW];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int ox0 = blockIdx.x * 32, oy0 = y0 + blockIdx.y * 8;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0)
    {
        mbar_arrive_expect_tx(&bar, uint32_t(sizeof(tile)));
        tma_load_2d(&tile[0][0], &in_map, 2 * (2 * ox0 - 2), 2 * oy0 - 2, &bar); // x in 64-bit elements: two per texel
    }
    mbar_wait(&bar, 0);
    bloom_down2x_from_tile<PREFILTER>(A, tile, corner, out, ox0, oy0, y1);
}
#endif

// B3 / B4 on an exact 1:2 level: the 3x3 tent of bilinear taps of the coarser level collapses to a separable 4-tap filter
// whose weights depend only on the parity of the output coordinate:
//   even x = 2k : texels k-2..k+1 weigh (1, 5, 7, 3)/16      odd x = 2k+1 : texels k-1..k+2 weigh (3, 7, 5, 1)/16
// (position x/2 - 1/4 resp. + 1/4 -> bilinear {1/4, 3/4}, convolved with the tent {1/4, 1/2, 1/4}). Clamp addressing is
// applied when the 20x8 coarse tile is staged. COMPOSITE selects B4 (lerp with Intensity) instead of B3 (plain add).
DFX_HD float3 tone_map_rt(int mode, float3 color, const dfx_tonemap_attribs& A, float aveLogLum);
DFX_HD float3 linear_to_srgb(float3 c);
struct ToneMapIn // only read by the TONEMAP variant: the final ToneMap(+sRGB) pass fused into the Bloom composite
{
    dfx_tonemap_attribs attribs;
    float               ave_log_lum;
    int                 to_srgb;
};

template <bool COMPOSITE, bool TONEMAP = false>
__global__ void __launch_bounds__(256) bloom_up2x_kernel(dfx_bloom_attribs A, View<const float4> fine, View<const float4> coarser, View<float4> out, int y0, int y1,
                                                         ToneMapIn tm = ToneMapIn{})
{
    // A CTA of 256 threads produces 64x16 outputs; every thread a 2x2 block that shares one 5x5 coarse footprint, so the
    // shared-memory traffic is 25 LDS.128 per four outputs (the kernel would otherwise be bound by smem bandwidth, not HBM).
    __shared__ float4 tile[12][36];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int ox0 = blockIdx.x * 64, oy0 = y0 + blockIdx.y * 16; // y0 is even for whole-level launches (checked by the caller)
    const int cx0 = (ox0 >> 1) - 2, cy0 = (oy0 >> 1) - 2;
    for (int i = tid; i < 12 * 36; i += 256)
    {
        const int ly = i / 36, lx = i - ly * 36;
        tile[ly][lx] = loadc(coarser, cx0 + lx, cy0 + ly);
    }
    __syncthreads();
    const int x = ox0 + 2 * threadIdx.x, y = oy0 + 2 * threadIdx.y;
    if (x >= out.w || y >= y1) return;
    float3 E[5], O[5]; // per coarse row: the horizontal 4-tap result for the even / odd output column
#pragma unroll
    for (int j = 0; j < 5; ++j)
    {
        const float3 c0 = xyz(tile[threadIdx.y + j][threadIdx.x]), c1 = xyz(tile[threadIdx.y + j][threadIdx.x + 1]), c2 = xyz(tile[threadIdx.y + j][threadIdx.x + 2]);
        const float3 c3 = xyz(tile[threadIdx.y + j][threadIdx.x + 3]), c4 = xyz(tile[threadIdx.y + j][threadIdx.x + 4]);
        E[j] = c0 * (1.f / 16) + c1 * (5.f / 16) + c2 * (7.f / 16) + c3 * (3.f / 16);
        O[j] = c1 * (3.f / 16) + c2 * (7.f / 16) + c3 * (5.f / 16) + c4 * (1.f / 16);
    }
    const float3 s00 = E[0] * (1.f / 16) + E[1] * (5.f / 16) + E[2] * (7.f / 16) + E[3] * (3.f / 16);
    const float3 s10 = O[0] * (1.f / 16) + O[1] * (5.f / 16) + O[2] * (7.f / 16) + O[3] * (3.f / 16);
    const float3 s01 = E[1] * (3.f / 16) + E[2] * (7.f / 16) + E[3] * (5.f / 16) + E[4] * (1.f / 16);
    const float3 s11 = O[1] * (3.f / 16) + O[2] * (7.f / 16) + O[3] * (5.f / 16) + O[4] * (1.f / 16);
    auto emit = [&](int px, int py, float3 s) {
        if (px >= out.w || py >= y1) return;
        const float3 c = xyz(__ldg(&fine.at(px, py))); // linear sampler at the texel centre == the texel
        if (COMPOSITE)
        {
            float3 o = lerp3(c, c + A.Intensity * s, A.AlphaInterpolation);
            if (TONEMAP)
            {
                o = tone_map_rt(tm.attribs.iToneMappingMode, o, tm.attribs, tm.ave_log_lum);
                if (tm.to_srgb) o = linear_to_srgb(o);
            }
            st_cs(&out.at(px, py), f4(o, 0.0f));
        }
        else
            out.at(px, py) = f4(c + s, 0.0f);
    };
    emit(x, y, s00), emit(x + 1, y, s10), emit(x, y + 1, s01), emit(x + 1, y + 1, s11);
}

// =====================================================================================================================
// TAA — TAA_ComputeTemporalAccumulation.fx:229-261
// =====================================================================================================================
template <bool YCOCG>
DFX_HD float3 rgb_to_ycocg(float3 c)
{
    if (!YCOCG) return c;
    float co = c.x - c.z, tmp = c.z + 0.5f * co, cg = c.y - tmp, yy = tmp + 0.5f * cg;
    return make_float3(yy, co, cg);
}
template <bool YCOCG>
DFX_HD float3 ycocg_to_rgb(float3 c)
{
    if (!YCOCG) return c;
    float tmp = c.x - 0.5f * c.z, g = c.z + tmp, b = tmp - 0.5f * c.y, r = b + c.y;
    return make_float3(r, g, b);
}
DFX_HD float3 hdr_to_sdr(float3 c) { return make_float3(c.x * frcp(1.0f + c.x), c.y * frcp(1.0f + c.y), c.z * frcp(1.0f + c.z)); }
DFX_HD float3 sdr_to_hdr(float3 c)
{
    return make_float3(c.x * frcp(1.0f - c.x + kFltEps), c.y * frcp(1.0f - c.y + kFltEps), c.z * frcp(1.0f - c.z + kFltEps));
}
DFX_HD float3 max0(float3 c) { return make_float3(fmaxf(c.x, 0.f), fmaxf(c.y, 0.f), fmaxf(c.z, 0.f)); }
DFX_HD float4 max0(float4 c) { return make_float4(fmaxf(c.x, 0.f), fmaxf(c.y, 0.f), fmaxf(c.z, 0.f), fmaxf(c.w, 0.f)); }

// ClipToAABB :98-106. Less/GreaterEqual are 0/1 selectors fed to lerp(a,b,t) = a + t*(b-a); fminf ignores NaN operands.
DFX_HD float3 clip_to_aabb(float3 prev, float3 curr, float3 centre, float3 ext)
{
    const float  maxT = 10.0f;
    const float3 dir  = curr - prev;
    const float  ix = fdiv((centre.x - signf(dir.x) * ext.x) - prev.x, dir.x); // x/0 -> +-inf, 0/0 -> NaN as in IEEE division
    const float  iy = fdiv((centre.y - signf(dir.y) * ext.y) - prev.y, dir.y);
    const float  iz = fdiv((centre.z - signf(dir.z) * ext.z) - prev.z, dir.z);
    const float  px = lerpf(maxT + 1.0f, ix, ix >= 0.0f ? 1.0f : 0.0f);
    const float  py = lerpf(maxT + 1.0f, iy, iy >= 0.0f ? 1.0f : 0.0f);
    const float  pz = lerpf(maxT + 1.0f, iz, iz >= 0.0f ? 1.0f : 0.0f);
    const float  T  = fminf(maxT, fminf(px, fminf(py, pz)));
    const float  lt = T < maxT ? 1.0f : 0.0f;
    return lerp3(prev, prev + dir * T, lt);
}

struct TaaCam
{
    CamS c, p;
};

// The compose step (rgb += ssr.rgb * ssr.a * scale; rgb *= lerp(1, ao, scale)) evaluated where the composed colour is
// consumed, so that the composed frame never makes a round trip through HBM (same arithmetic as compose_kernel).
struct ComposeIn
{
    View<const float4> ssr;
    View<const float>  ao;
    float              ssr_scale, ssao_scale;
};
template <bool COMPOSE>
DFX_HD float3 load_scene_colour(const View<const float4>& color, const ComposeIn& ci, int gx, int gy)
{
    float3 c = xyz(__ldg(&color.at(gx, gy)));
    if (COMPOSE)
    {
        if (ci.ssr.p && ci.ssr_scale > 0.0f)
        {
            const float4 s = __ldg(&ci.ssr.at(gx, gy));
            c              = c + xyz(s) * s.w * ci.ssr_scale;
        }
        if (ci.ao.p && ci.ssao_scale > 0.0f) c = c * lerpf(1.0f, __ldg(&ci.ao.at(gx, gy)), ci.ssao_scale);
    }
    return c;
}

template <bool BICUBIC, bool YCOCG, bool GAUSS, bool COMPOSE>
__global__ void __launch_bounds__(256, DFX_OCC_TAA) taa_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_taa_attribs A, View<const float4> curr_color,
                                                  ComposeIn ci, View<const float4> prev_accum, View<const float2> motion, View<const float> curr_depth,
                                                  View<const float> prev_depth, View<float4> out, int y0, int y1)
{
    // 32x8 pixel tile + 1-pixel halo of the current colour, converted ONCE per texel to the clipping space (Reinhard SDR,
    // optionally YCoCg) and shared through smem: the 3x3 statistics then cost 9 LDS instead of 9 LDG + 9 conversions.
    __shared__ TaaCam S;
    __shared__ float4 tile[10][34];
    if (threadIdx.x == 0 && threadIdx.y == 0) load_cam(S.c, &cams[0]), load_cam(S.p, &cams[1]);
    {
        const int tx0 = blockIdx.x * 32 - 1, ty0 = y0 + blockIdx.y * 8 - 1;
        for (int i = threadIdx.y * 32 + threadIdx.x; i < 340; i += 256)
        {
            const int    ly = i / 34, lx = i - ly * 34;
            const int    gx = min(max(tx0 + lx, 0), curr_color.w - 1), gy = min(max(ty0 + ly, 0), curr_color.h - 1); // ClampScreenCoord
            const float3 sdr = rgb_to_ycocg<YCOCG>(hdr_to_sdr(max0(load_scene_colour<COMPOSE>(curr_color, ci, gx, gy))));
            tile[ly][lx]     = f4(sdr, 0.0f);
        }
    }
    __syncthreads();
    const CamS& cam = S.c;
    const int   x = blockIdx.x * blockDim.x + threadIdx.x;
    const int   y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out.w || y >= y1) return;

    const float posx = float(x) + 0.5f, posy = float(y) + 0.5f;
    float2      mv   = __ldg(&motion.at(x, y));
    mv.x *= 0.5f, mv.y *= -0.5f;
    const float ppx = posx - mv.x * cam.vw, ppy = posy - mv.y * cam.vh;
    const float3 currHDR = max0(load_scene_colour<COMPOSE>(curr_color, ci, x, y));

    if (!(ppx >= 0.0f && ppy >= 0.0f && ppx < cam.vw && ppy < cam.vh) || A.ResetAccumulation)
    {
        st_cs(&out.at(x, y), f4(currHDR, 0.5f));
        return;
    }
    const float aspect = cam.vw * cam.ivh;
    const float mf     = saturate(1.0f - fsqrt((mv.x * aspect) * (mv.x * aspect) + mv.y * mv.y) * 256.0f);

    // ComputeDepthDisocclusion :117-136 (unclamped loads)
    float depthFactor;
    {
        const int   pix = (int)ppx, piy = (int)ppy;
        const float cd  = __ldg(&curr_depth.at(x, y));
        const float lc  = fabsf(depth_to_camz(cd, cam));
        // max_i exp(-r_i) > 0.9  <=>  min_i r_i < -ln(0.9): the predicate is evaluated without the nine exp()
        float rmin = kFltMax;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx)
            {
                const float pd = load0(prev_depth, pix + dx, piy + dy);
                const float lp = fabsf(fdiv(S.p.m32 - pd * S.p.m33, pd * S.p.m23 - S.p.m22));
                rmin           = fminf(rmin, fdiv(fabsf(lc - lp), fmaxf(fmaxf(lc, lp), 1e-6f)));
            }
        depthFactor = rmin < 0.105360516f ? 1.0f : 0.0f;
    }

    float4 prevHDR;
    if (BICUBIC)
    {
        // SamplePrevColorCatmullRom :138-173 (5 bilinear taps)
        const float cx = floorf(ppx - 0.5f) + 0.5f, cy = floorf(ppy - 0.5f) + 0.5f;
        const float fx = ppx - cx, fy = ppy - cy;
        const float fx2 = fx * fx, fy2 = fy * fy, fx3 = fx2 * fx, fy3 = fy2 * fy;
        const float w0x = -0.5f * fx3 + fx2 - 0.5f * fx, w0y = -0.5f * fy3 + fy2 - 0.5f * fy;
        const float w1x = 1.5f * fx3 - 2.5f * fx2 + 1.0f, w1y = 1.5f * fy3 - 2.5f * fy2 + 1.0f;
        const float w2x = -1.5f * fx3 + 2.0f * fx2 + 0.5f * fx, w2y = -1.5f * fy3 + 2.0f * fy2 + 0.5f * fy;
        const float w3x = 0.5f * fx3 - 0.5f * fx2, w3y = 0.5f * fy3 - 0.5f * fy2;
        const float w12x = w1x + w2x, w12y = w1y + w2y;
        const float t0x = (cx - 1.0f) * cam.ivw, t0y = (cy - 1.0f) * cam.ivh;
        const float t3x = (cx + 2.0f) * cam.ivw, t3y = (cy + 2.0f) * cam.ivh;
        const float t12x = (cx + fdiv(w2x, w12x)) * cam.ivw, t12y = (cy + fdiv(w2y, w12y)) * cam.ivh;
        const float p0 = w12x * w0y, p1 = w0x * w12y, p2 = w12x * w12y, p3 = w3x * w12y, p4 = w12x * w3y;
        // The five bilinear taps touch 12 texels, not 20: the "0" and "3" coordinates are exact texel centres (after the
        // sampler's 1/256 snap their second bilinear weight is exactly 0), only the "12" coordinate blends two texels.
        const int   PW = prev_accum.w, PH = prev_accum.h;
        const float fpw = float(PW), fph = float(PH);
        const float sx12 = snap8(t12x * fpw - 0.5f), sy12 = snap8(t12y * fph - 0.5f);
        const float bx = floorf(sx12), by = floorf(sy12);
        const float qx = sx12 - bx, qy = sy12 - by; // weights of the right / lower texel of the "12" pair
        auto cxi = [&](int v) { return min(max(v, 0), PW - 1); };
        auto cyi = [&](int v) { return min(max(v, 0), PH - 1); };
        const int xa = cxi((int)bx), xb = cxi((int)bx + 1), ya = cyi((int)by), yb = cyi((int)by + 1);
        const int x0 = cxi((int)rintf(snap8(t0x * fpw - 0.5f))), x3 = cxi((int)rintf(snap8(t3x * fpw - 0.5f)));
        const int y0i = cyi((int)rintf(snap8(t0y * fph - 0.5f))), y3i = cyi((int)rintf(snap8(t3y * fph - 0.5f)));
        auto ld = [&](int tx, int ty) { return __ldg(&prev_accum.at(tx, ty)); };
        auto mixx = [&](int ty) { return ld(xa, ty) * (1.0f - qx) + ld(xb, ty) * qx; };
        const float4 rowa = mixx(ya), rowb = mixx(yb);
        float4       r    = mixx(y0i) * p0;                                        // (12, 0)
        r = r + (ld(x0, ya) * (1.0f - qy) + ld(x0, yb) * qy) * p1;                 // (0, 12)
        r = r + (rowa * (1.0f - qy) + rowb * qy) * p2;                             // (12, 12)
        r = r + (ld(x3, ya) * (1.0f - qy) + ld(x3, yb) * qy) * p3;                 // (3, 12)
        r = r + mixx(y3i) * p4;                                                    // (12, 3)
        prevHDR = max0(r * frcp(p0 + p1 + p2 + p3 + p4));
    }
    else
    {
        prevHDR = max0(sample_linear_clamp(prev_accum, ppx * cam.ivw, ppy * cam.ivh));
    }

    const float3 currSDR = rgb_to_ycocg<YCOCG>(hdr_to_sdr(currHDR));
    const float3 prevSDR = rgb_to_ycocg<YCOCG>(hdr_to_sdr(xyz(prevHDR)));
    auto corrected_alpha = [&](float a) { return fminf(A.TemporalStabilityFactor, saturate(frcp(2.0f - a))); };

    if (A.SkipRejection)
    {
        const float3 o = sdr_to_hdr(ycocg_to_rgb<YCOCG>(lerp3(currSDR, prevSDR, prevHDR.w)));
        st_cs(&out.at(x, y), f4(o, corrected_alpha(prevHDR.w)));
        return;
    }

    const float gamma = lerpf(0.75f, 2.5f, mf * mf);
    // ComputePixelStatisticYCoCgSDR :191-222
    float3 m1 = make_float3(0.f, 0.f, 0.f), m2 = m1;
    float  wsum = 0.0f;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
        {
            const float3 sdr = xyz(tile[threadIdx.y + 1 + dy][threadIdx.x + 1 + dx]);
            // exp(-3 r^2 / 4), r^2 in {0, 1, 2}: compile-time constants after unrolling
            const float  w   = GAUSS ? ((dx * dx + dy * dy) == 0 ? 1.0f : (dx * dx + dy * dy) == 1 ? 0.472366553f : 0.223130160f) : 1.0f;
            m1 = m1 + sdr * w, m2 = m2 + sdr * sdr * w;
            wsum += w;
        }
    const float3 mean = m1 * frcp(wsum);
    const float3 var  = m2 * frcp(wsum) - (mean * mean);
    const float3 sd   = make_float3(fsqrt(fmaxf(var.x, 0.f)), fsqrt(fmaxf(var.y, 0.f)), fsqrt(fmaxf(var.z, 0.f)));
    const float3 clipped = clip_to_aabb(prevSDR, currSDR, mean, gamma * sd);
    const float  alpha   = prevHDR.w * mf * depthFactor;
    const float3 o       = sdr_to_hdr(ycocg_to_rgb<YCOCG>(lerp3(currSDR, clipped, alpha)));
    st_cs(&out.at(x, y), f4(o, corrected_alpha(alpha)));
}

// =====================================================================================================================
// compose (reduced form of HnPostProcess.psh:145-185): rgb += ssr.rgb*ssr.a*scale ; rgb *= lerp(1, ao, scale)
// =====================================================================================================================
__global__ void __launch_bounds__(256) compose_kernel(View<const float4> color, View<const float4> ssr, View<const float> ao, float ssr_scale,
                                                      float ssao_scale, View<float4> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float4 C = __ldg(&color.at(x, y));
    float3       c = xyz(C);
    if (ssr.p && ssr_scale > 0.0f)
    {
        const float4 s = __ldg(&ssr.at(x, y));
        c              = c + xyz(s) * s.w * ssr_scale;
    }
    if (ao.p && ssao_scale > 0.0f) c = c * lerpf(1.0f, __ldg(&ao.at(x, y)), ssao_scale);
    st_cs(&out.at(x, y), f4(c, C.w));
}

// =====================================================================================================================
// ToneMap — ToneMapping.fxh:87-226 (all 11 operators) + LinearToSRGB (SRGBUtilities.fxh:27-33)
// =====================================================================================================================
DFX_HD float3 uncharted2(float3 x) // :8-19
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    auto        f = [&](float v) { return fdiv(v * (A * v + C * B) + D * E, v * (A * v + B) + D * F) - E / F; };
    return make_float3(f(x.x), f(x.y), f(x.z));
}
// pow via MUFU lg2/ex2 (relative error ~1e-6 at these exponents): the tone map is a full-screen element-wise pass that
// the correctly-rounded powf() would make ALU-bound instead of HBM-bound
DFX_HD float fpow(float v, float e) { return __powf(v, e); }
DFX_HD float3 pow3(float3 v, float e) { return make_float3(fpow(v.x, e), fpow(v.y, e), fpow(v.z, e)); }
DFX_HD float3 srgb_to_linear(float3 s)
{
    auto f = [](float v) {
        float hi = fpow(saturate((v + 0.055f) * (1.0f / 1.055f)), 2.4f);
        return lerpf(v / 12.92f, hi, v >= 0.04045f ? 1.0f : 0.0f);
    };
    return make_float3(f(s.x), f(s.y), f(s.z));
}
DFX_HD float3 linear_to_srgb(float3 c)
{
    auto f = [](float v) {
        float hi = fpow(v, 1.0f / 2.4f) * 1.055f - 0.055f;
        return lerpf(v * 12.92f, hi, v >= 0.0031308f ? 1.0f : 0.0f);
    };
    return make_float3(f(c.x), f(c.y), f(c.z));
}
DFX_HD float3 agx(float3 c) // :36-58
{
    float3 t = make_float3(0.842479062253094f * c.x + 0.0784335999999992f * c.y + 0.0792237451477643f * c.z,
                           0.0423282422610123f * c.x + 0.878468636469772f * c.y + 0.0791661274605434f * c.z,
                           0.0423756549057051f * c.x + 0.0784336f * c.y + 0.879142973793104f * c.z);
    const float mn = -12.47393f, mx = 4.026069f;
    auto        enc = [&](float v) { return (fminf(fmaxf(log2f(v), mn), mx) - mn) / (mx - mn); };
    t               = make_float3(enc(t.x), enc(t.y), enc(t.z));
    auto poly = [](float x) {
        float x2 = x * x, x4 = x2 * x2;
        return 15.5f * x4 * x2 - 40.14f * x4 * x + 31.96f * x4 - 6.868f * x2 * x + 0.4298f * x2 + 0.1191f * x - 0.00232f;
    };
    return make_float3(poly(t.x), poly(t.y), poly(t.z));
}
DFX_HD float3 agx_eotf(float3 c) // :60-74
{
    float3 t = make_float3(+1.19687900512017f * c.x - 0.0980208811401368f * c.y - 0.0990297440797205f * c.z,
                           -0.0528968517574562f * c.x + 1.15190312990417f * c.y - 0.0989611768448433f * c.z,
                           -0.0529716355144438f * c.x - 0.0980434501171241f * c.y + 1.15107367264116f * c.z);
    return srgb_to_linear(t);
}

template <int MODE>
DFX_HD float3 tone_map(float3 color, const dfx_tonemap_attribs& A, float aveLogLum)
{
    const float3 lumw  = make_float3(0.212671f, 0.715160f, 0.072169f);
    const float  scale = A.fMiddleGray / aveLogLum;
    color              = max0(color);
    const float  lum0  = fmaxf(dot(lumw, color), 1e-10f);
    const float  lumS  = lum0 * scale;
    const float3 cS    = color * scale;
    const float  wp    = A.fWhitePoint;
    if (MODE == DFX_TONE_MAPPING_MODE_EXP) return (1.0f - expf(-lumS)) * pow3(color / lum0, A.fLuminanceSaturation);
    if (MODE == DFX_TONE_MAPPING_MODE_REINHARD) return (lumS / (1.0f + lumS)) * pow3(color / lum0, A.fLuminanceSaturation);
    if (MODE == DFX_TONE_MAPPING_MODE_REINHARD_MOD) return (lumS * (1.0f + lumS / (wp * wp)) / (1.0f + lumS)) * pow3(color / lum0, A.fLuminanceSaturation);
    if (MODE == DFX_TONE_MAPPING_MODE_UNCHARTED2)
    {
        const float3 curr = uncharted2(2.0f * cS);
        const float3 w    = uncharted2(make_float3(wp, wp, wp));
        return curr * make_float3(frcp(w.x), frcp(w.y), frcp(w.z));
    }
    if (MODE == DFX_TONE_MAPPING_MODE_FILMIC_ALU)
    {
        auto f = [](float v) {
            v = fmaxf(v - 0.004f, 0.0f);
            v = (v * (6.2f * v + 0.5f)) / (v * (6.2f * v + 1.7f) + 0.06f);
            return powf(v, 2.2f);
        };
        return make_float3(f(cS.x), f(cS.y), f(cS.z));
    }
    if (MODE == DFX_TONE_MAPPING_MODE_LOGARITHMIC) return (log10f(1.0f + lumS) / log10f(1.0f + wp)) * pow3(color / lum0, A.fLuminanceSaturation);
    if (MODE == DFX_TONE_MAPPING_MODE_ADAPTIVE_LOG)
    {
        const float l = 1.0f / log10f(1.0f + wp) * logf(1.0f + lumS) / logf(2.0f + 8.0f * powf(lumS / wp, logf(0.85f) / logf(0.5f)));
        return l * pow3(color / lum0, A.fLuminanceSaturation);
    }
    if (MODE == DFX_TONE_MAPPING_MODE_AGX) return agx_eotf(agx(cS));
    if (MODE == DFX_TONE_MAPPING_MODE_AGX_CUSTOM)
    {
        float3      c   = agx(cS);
        const float lum = dot(c, lumw);
        c               = pow3(c * A.AgXSlope + make_float3(A.AgXOffset, A.AgXOffset, A.AgXOffset), A.AgXPower);
        c               = make_float3(lum, lum, lum) + A.AgXSaturation * (c - make_float3(lum, lum, lum));
        return agx_eotf(c);
    }
    if (MODE == DFX_TONE_MAPPING_MODE_PBR_NEUTRAL)
    {
        float3      c   = color * (0.3f / aveLogLum);
        const float sc  = 0.8f - 0.04f, desat = 0.15f;
        const float mn  = fminf(c.x, fminf(c.y, c.z));
        const float off = mn < 0.08f ? mn - 6.25f * mn * mn : 0.04f;
        c               = c - make_float3(off, off, off);
        const float peak = fmaxf(c.x, fmaxf(c.y, c.z));
        if (peak >= sc)
        {
            const float d = 1.0f - sc, np = 1.0f - d * d / (peak + d - sc);
            c             = c * (np / peak);
            const float g = 1.0f - 1.0f / (desat * (peak - np) + 1.0f);
            c             = lerp3(c, make_float3(np, np, np), g);
        }
        return c;
    }
    if (MODE == DFX_TONE_MAPPING_MODE_COMMERCE)
    {
        float3      c  = color * (0.3f / aveLogLum);
        const float sc = 0.8f, desat = 0.5f, d = 1.0f - sc;
        const float peak = fmaxf(c.x, fmaxf(c.y, c.z));
        if (peak >= sc)
        {
            const float np = 1.0f - d * d / (peak + d - sc), ip = 1.0f / peak;
            const float3 e = c * (1.0f - sc * ip);
            const float extra = e.x * 1.0f + e.y * 1.0f + e.z * 1.0f;
            c             = c * (np * ip);
            const float g = 1.0f - 3.0f / (desat * extra + 3.0f);
            c             = lerp3(c, make_float3(1.f, 1.f, 1.f), g);
        }
        return c;
    }
    return color;
}

DFX_HD float3 tone_map_rt(int mode, float3 color, const dfx_tonemap_attribs& A, float aveLogLum)
{
    switch (mode) // warp-uniform
    {
        case DFX_TONE_MAPPING_MODE_EXP: return tone_map<DFX_TONE_MAPPING_MODE_EXP>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_REINHARD: return tone_map<DFX_TONE_MAPPING_MODE_REINHARD>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_REINHARD_MOD: return tone_map<DFX_TONE_MAPPING_MODE_REINHARD_MOD>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_UNCHARTED2: return tone_map<DFX_TONE_MAPPING_MODE_UNCHARTED2>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_FILMIC_ALU: return tone_map<DFX_TONE_MAPPING_MODE_FILMIC_ALU>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_LOGARITHMIC: return tone_map<DFX_TONE_MAPPING_MODE_LOGARITHMIC>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_ADAPTIVE_LOG: return tone_map<DFX_TONE_MAPPING_MODE_ADAPTIVE_LOG>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_AGX: return tone_map<DFX_TONE_MAPPING_MODE_AGX>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_AGX_CUSTOM: return tone_map<DFX_TONE_MAPPING_MODE_AGX_CUSTOM>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_PBR_NEUTRAL: return tone_map<DFX_TONE_MAPPING_MODE_PBR_NEUTRAL>(color, A, aveLogLum);
        case DFX_TONE_MAPPING_MODE_COMMERCE: return tone_map<DFX_TONE_MAPPING_MODE_COMMERCE>(color, A, aveLogLum);
        default: return tone_map<DFX_TONE_MAPPING_MODE_NONE>(color, A, aveLogLum);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) tonemap_kernel(dfx_tonemap_attribs A, float aveLogLum, int to_srgb, View<const float4> in, View<float4> out,
                                                      int y0, int y1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = y0 + blockIdx.y;
    if (x >= out.w || y >= y1) return;
    const float4 C = __ldg(&in.at(x, y));
    float3       c = tone_map<MODE>(xyz(C), A, aveLogLum);
    if (to_srgb) c = linear_to_srgb(c);
    st_cs(&out.at(x, y), f4(c, C.w));
}

} // namespace dfx

using namespace dfx;

#define DFX_GRID(w, rows) dim3 block(32, 8), grid(div_up(w, 32), div_up(rows.y1 - rows.y0, 8))

#if DFX_BLOOM_TMA
#    include <map>
#    include <mutex>
#    include <tuple>
// Tensor maps of the planes the TMA variant has read so far (encoding one costs a driver call; a pyramid has a dozen planes).
static const CUtensorMap* bloom_source_map(const View<const float4>& in)
{
    static std::mutex                                                              m;
    static std::map<std::tuple<const void*, int, int, int>, CUtensorMap>           cache;
    std::lock_guard<std::mutex>                                                    lk(m);
    const auto                                                                     key = std::make_tuple(static_cast<const void*>(in.p), in.w, in.h, in.pitch);
    auto                                                                           it  = cache.find(key);
    if (it != cache.end()) return &it->second;
    CUtensorMap map;
    if (!make_tensor_map_rgba32f(&map, in.p, in.w, in.h, size_t(in.pitch) * sizeof(float4), kDnTileW, kDnTileH)) return nullptr;
    return &cache.emplace(key, map).first->second;
}
#endif

static inline dfx_rows scale_rows(dfx_rows r, int full_h, int h)
{
    // rows of a plane of height h that correspond to the full-frame strip r (h = full_h >> k)
    if (full_h == h) return r;
    int k = 0;
    while ((full_h >> k) > h && k < 16) ++k;
    dfx_rows o;
    o.y0 = r.y0 >> k;
    o.y1 = r.y1 >= full_h ? h : (r.y1 >> k);
    return o;
}

extern "C" dfx_status dfx_pass_bloom_prefilter(void* stream, const dfx_bloom_attribs* attribs, const dfx_plane* color, const dfx_plane* out_level0, dfx_rows rows)
{
    DFX_PROFILE(stream, "bloom_prefilter");
    DFX_REQUIRE(attribs, "null argument");
    DFX_VIEW(const float4, in, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, out, out_level0, DFX_FORMAT_RGBA32F);
    DFX_REQUIRE(out.w == max(in.w / 2, 1) && out.h == max(in.h / 2, 1), "level 0 must be half the input size");
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range (rows are in output-plane coordinates)");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(out.w, rows);
#if DFX_BLOOM_TMA
    if (const CUtensorMap* map = (in.w == 2 * out.w && in.h == 2 * out.h) ? bloom_source_map(in) : nullptr)
        bloom_down2x_tma_kernel<true><<<grid, block, 0, as_stream(stream)>>>(*attribs, *map, out, rows.y0, rows.y1);
    else
#endif
    if (in.w == 2 * out.w && in.h == 2 * out.h)
        bloom_down2x_kernel<true><<<grid, block, 0, as_stream(stream)>>>(*attribs, in, out, rows.y0, rows.y1);
    else
        bloom_prefilter_kernel<<<grid, block, 0, as_stream(stream)>>>(*attribs, in, out, rows.y0, rows.y1);
    DFX_LAUNCHED("bloom_prefilter_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_bloom_downsample(void* stream, const dfx_plane* in_, const dfx_plane* out_, dfx_rows rows)
{
    DFX_PROFILE(stream, "bloom_downsample");
    DFX_VIEW(const float4, in, in_, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, out, out_, DFX_FORMAT_RGBA32F);
    DFX_REQUIRE(out.w == max(in.w / 2, 1) && out.h == max(in.h / 2, 1), "output must be half the input size");
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range (rows are in output-plane coordinates)");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(out.w, rows);
#if DFX_BLOOM_TMA
    if (const CUtensorMap* map = (in.w == 2 * out.w && in.h == 2 * out.h) ? bloom_source_map(in) : nullptr)
        bloom_down2x_tma_kernel<false><<<grid, block, 0, as_stream(stream)>>>(dfx_bloom_attribs{}, *map, out, rows.y0, rows.y1);
    else
#endif
    if (in.w == 2 * out.w && in.h == 2 * out.h)
        bloom_down2x_kernel<false><<<grid, block, 0, as_stream(stream)>>>(dfx_bloom_attribs{}, in, out, rows.y0, rows.y1);
    else
        bloom_downsample_kernel<<<grid, block, 0, as_stream(stream)>>>(in, out, rows.y0, rows.y1);
    DFX_LAUNCHED("bloom_downsample_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_bloom_upsample(void* stream, const dfx_plane* same_level_down, const dfx_plane* coarser, const dfx_plane* out_, dfx_rows rows)
{
    DFX_PROFILE(stream, "bloom_upsample");
    DFX_VIEW(const float4, same, same_level_down, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, lo, coarser, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, out, out_, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(same, out);
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range (rows are in output-plane coordinates)");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(out.w, rows);
    if (out.w == 2 * lo.w && out.h == 2 * lo.h && (rows.y0 & 1) == 0)
        bloom_up2x_kernel<false><<<dim3(div_up(out.w, 64), div_up(rows.y1 - rows.y0, 16)), block, 0, as_stream(stream)>>>(dfx_bloom_attribs{}, same, lo, out, rows.y0, rows.y1);
    else
        bloom_upsample_kernel<<<grid, block, 0, as_stream(stream)>>>(same, lo, out, rows.y0, rows.y1);
    DFX_LAUNCHED("bloom_upsample_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_bloom_composite(void* stream, const dfx_bloom_attribs* attribs, const dfx_plane* color, const dfx_plane* up0,
                                               const dfx_plane* out_, dfx_rows rows)
{
    DFX_PROFILE(stream, "bloom_composite");
    DFX_REQUIRE(attribs, "null argument");
    DFX_VIEW(const float4, c, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, u, up0, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, out, out_, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(c, out);
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(out.w, rows);
    if (out.w == 2 * u.w && out.h == 2 * u.h && (rows.y0 & 1) == 0)
        bloom_up2x_kernel<true><<<dim3(div_up(out.w, 64), div_up(rows.y1 - rows.y0, 16)), block, 0, as_stream(stream)>>>(*attribs, c, u, out, rows.y0, rows.y1);
    else
        bloom_composite_kernel<<<grid, block, 0, as_stream(stream)>>>(*attribs, c, u, out, rows.y0, rows.y1);
    DFX_LAUNCHED("bloom_composite_kernel");
    return DFX_OK;
}

// B4 + M1/M2 in one kernel (only on exact 2:1 levels: the caller falls back to the two separate passes otherwise).
extern "C" dfx_status dfx_pass_bloom_composite_tonemap(void* stream, const dfx_bloom_attribs* attribs, const dfx_tonemap_attribs* tonemap, float ave_log_lum,
                                                       int32_t convert_to_srgb, const dfx_plane* color, const dfx_plane* up0, const dfx_plane* ldr_out, dfx_rows rows)
{
    DFX_PROFILE(stream, "bloom_composite_tonemap");
    DFX_REQUIRE(attribs && tonemap, "null argument");
    DFX_VIEW(const float4, c, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, u, up0, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, out, ldr_out, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(c, out);
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (!(out.w == 2 * u.w && out.h == 2 * u.h && (rows.y0 & 1) == 0)) return set_error(DFX_ERR_UNSUPPORTED, "fused composite+tonemap needs an exact 2:1 level");
    DFX_REQUIRE(tonemap->iToneMappingMode >= 0 && tonemap->iToneMappingMode <= DFX_TONE_MAPPING_MODE_COMMERCE, "unknown tone mapping mode %d", tonemap->iToneMappingMode);
    if (rows.y1 == rows.y0) return DFX_OK;
    bloom_up2x_kernel<true, true><<<dim3(div_up(out.w, 64), div_up(rows.y1 - rows.y0, 16)), dim3(32, 8), 0, as_stream(stream)>>>(
        *attribs, c, u, out, rows.y0, rows.y1, ToneMapIn{*tonemap, ave_log_lum, convert_to_srgb});
    DFX_LAUNCHED("bloom_up2x_kernel<composite, tonemap>");
    return DFX_OK;
}

static dfx_status launch_taa(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_taa_attribs* attribs, uint32_t flags, bool compose,
                             const dfx_plane* ssr, const dfx_plane* ao, float ssr_scale, float ssao_scale, const dfx_plane* curr_color,
                             const dfx_plane* prev_accum, const dfx_plane* closest_motion, const dfx_plane* reprojected_depth,
                             const dfx_plane* previous_depth, const dfx_plane* out_accum, dfx_rows rows)
{
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const float4, cc, curr_color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, pa, prev_accum, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float2, mv, closest_motion, DFX_FORMAT_RG32F);
    DFX_VIEW(const float, cd, reprojected_depth, DFX_FORMAT_R32F);
    DFX_VIEW(const float, pd, previous_depth, DFX_FORMAT_R32F);
    DFX_VIEW(float4, out, out_accum, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(cc, pa);
    DFX_SAME_SIZE(cc, mv);
    DFX_SAME_SIZE(cc, cd);
    DFX_SAME_SIZE(cc, pd);
    DFX_SAME_SIZE(cc, out);
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    ComposeIn ci{View<const float4>{nullptr, 0, 0, 0}, View<const float>{nullptr, 0, 0, 0}, ssr_scale, ssao_scale};
    if (compose && ssr) DFX_REQUIRE(make_view<const float4>(ssr, DFX_FORMAT_RGBA32F, ci.ssr) && ci.ssr.w == cc.w && ci.ssr.h == cc.h, "bad ssr plane");
    if (compose && ao) DFX_REQUIRE(make_view<const float>(ao, DFX_FORMAT_R32F, ci.ao) && ci.ao.w == cc.w && ci.ao.h == cc.h, "bad ao plane");
    DFX_GRID(out.w, rows);
    cudaStream_t s = as_stream(stream);
#define TAA_LAUNCH(B, Y, G)                                                                                                            \
    do {                                                                                                                               \
        if (compose)                                                                                                                   \
            taa_kernel<B, Y, G, true><<<grid, block, 0, s>>>(cameras_dev, *attribs, cc, ci, pa, mv, cd, pd, out, rows.y0, rows.y1);    \
        else                                                                                                                           \
            taa_kernel<B, Y, G, false><<<grid, block, 0, s>>>(cameras_dev, *attribs, cc, ci, pa, mv, cd, pd, out, rows.y0, rows.y1);   \
    } while (0)
    switch (flags & 7u)
    {
        case 0: TAA_LAUNCH(false, false, false); break;
        case 1: TAA_LAUNCH(false, false, true); break;
        case 2: TAA_LAUNCH(true, false, false); break;
        case 3: TAA_LAUNCH(true, false, true); break;
        case 4: TAA_LAUNCH(false, true, false); break;
        case 5: TAA_LAUNCH(false, true, true); break;
        case 6: TAA_LAUNCH(true, true, false); break;
        case 7: TAA_LAUNCH(true, true, true); break;
    }
#undef TAA_LAUNCH
    DFX_LAUNCHED("taa_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_taa(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_taa_attribs* attribs, uint32_t flags,
                                   const dfx_plane* curr_color, const dfx_plane* prev_accum, const dfx_plane* closest_motion,
                                   const dfx_plane* reprojected_depth, const dfx_plane* previous_depth, const dfx_plane* out_accum, dfx_rows rows)
{
    DFX_PROFILE(stream, "taa");
    return launch_taa(stream, cameras_dev, attribs, flags, false, nullptr, nullptr, 0.0f, 0.0f, curr_color, prev_accum, closest_motion, reprojected_depth,
                      previous_depth, out_accum, rows);
}

extern "C" dfx_status dfx_pass_compose_taa(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_taa_attribs* attribs, uint32_t flags,
                                           const dfx_plane* color, const dfx_plane* ssr, const dfx_plane* ao, float ssr_scale, float ssao_scale,
                                           const dfx_plane* prev_accum, const dfx_plane* closest_motion, const dfx_plane* reprojected_depth,
                                           const dfx_plane* previous_depth, const dfx_plane* out_accum, dfx_rows rows)
{
    DFX_PROFILE(stream, "compose_taa");
    return launch_taa(stream, cameras_dev, attribs, flags, true, ssr, ao, ssr_scale, ssao_scale, color, prev_accum, closest_motion, reprojected_depth,
                      previous_depth, out_accum, rows);
}

extern "C" dfx_status dfx_pass_compose(void* stream, const dfx_plane* color, const dfx_plane* ssr, const dfx_plane* ao, float ssr_scale,
                                       float ssao_scale, const dfx_plane* out_, dfx_rows rows)
{
    DFX_PROFILE(stream, "compose");
    DFX_VIEW(const float4, c, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, out, out_, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(c, out);
    View<const float4> s{nullptr, 0, 0, 0};
    View<const float>  a{nullptr, 0, 0, 0};
    if (ssr)
    {
        DFX_REQUIRE(make_view<const float4>(ssr, DFX_FORMAT_RGBA32F, s) && s.w == c.w && s.h == c.h, "bad ssr plane");
    }
    if (ao)
    {
        DFX_REQUIRE(make_view<const float>(ao, DFX_FORMAT_R32F, a) && a.w == c.w && a.h == c.h, "bad ao plane");
    }
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(out.w, rows);
    compose_kernel<<<grid, block, 0, as_stream(stream)>>>(c, s, a, ssr_scale, ssao_scale, out, rows.y0, rows.y1);
    DFX_LAUNCHED("compose_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_tonemap(void* stream, const dfx_tonemap_attribs* attribs, float ave_log_lum, int32_t convert_to_srgb,
                                       const dfx_plane* color, const dfx_plane* out_, dfx_rows rows)
{
    DFX_PROFILE(stream, "tonemap");
    DFX_REQUIRE(attribs, "null argument");
    DFX_VIEW(const float4, c, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, out, out_, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(c, out);
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3         block(256), grid(div_up(out.w, 256), rows.y1 - rows.y0);
    cudaStream_t s = as_stream(stream);
#define TM_LAUNCH(M) tonemap_kernel<M><<<grid, block, 0, s>>>(*attribs, ave_log_lum, convert_to_srgb, c, out, rows.y0, rows.y1)
    switch (attribs->iToneMappingMode)
    {
        case DFX_TONE_MAPPING_MODE_NONE: TM_LAUNCH(DFX_TONE_MAPPING_MODE_NONE); break;
        case DFX_TONE_MAPPING_MODE_EXP: TM_LAUNCH(DFX_TONE_MAPPING_MODE_EXP); break;
        case DFX_TONE_MAPPING_MODE_REINHARD: TM_LAUNCH(DFX_TONE_MAPPING_MODE_REINHARD); break;
        case DFX_TONE_MAPPING_MODE_REINHARD_MOD: TM_LAUNCH(DFX_TONE_MAPPING_MODE_REINHARD_MOD); break;
        case DFX_TONE_MAPPING_MODE_UNCHARTED2: TM_LAUNCH(DFX_TONE_MAPPING_MODE_UNCHARTED2); break;
        case DFX_TONE_MAPPING_MODE_FILMIC_ALU: TM_LAUNCH(DFX_TONE_MAPPING_MODE_FILMIC_ALU); break;
        case DFX_TONE_MAPPING_MODE_LOGARITHMIC: TM_LAUNCH(DFX_TONE_MAPPING_MODE_LOGARITHMIC); break;
        case DFX_TONE_MAPPING_MODE_ADAPTIVE_LOG: TM_LAUNCH(DFX_TONE_MAPPING_MODE_ADAPTIVE_LOG); break;
        case DFX_TONE_MAPPING_MODE_AGX: TM_LAUNCH(DFX_TONE_MAPPING_MODE_AGX); break;
        case DFX_TONE_MAPPING_MODE_AGX_CUSTOM: TM_LAUNCH(DFX_TONE_MAPPING_MODE_AGX_CUSTOM); break;
        case DFX_TONE_MAPPING_MODE_PBR_NEUTRAL: TM_LAUNCH(DFX_TONE_MAPPING_MODE_PBR_NEUTRAL); break;
        case DFX_TONE_MAPPING_MODE_COMMERCE: TM_LAUNCH(DFX_TONE_MAPPING_MODE_COMMERCE); break;
        default: return set_error(DFX_ERR_INVALID_ARG, "unknown tone mapping mode %d", attribs->iToneMappingMode);
    }
#undef TM_LAUNCH
    DFX_LAUNCHED("tonemap_kernel");
    return DFX_OK;
}
