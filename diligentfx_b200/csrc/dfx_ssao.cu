// dfx_ssao.cu — ScreenSpaceAmbientOcclusion passes A1-A8 as sm_100a kernels.
// Reference host code: PostProcess/ScreenSpaceAmbientOcclusion/src/ScreenSpaceAmbientOcclusion.cpp:818-1329;
// shaders: Shaders/PostProcess/ScreenSpaceAmbientOcclusion/private/SSAO_*.fx (cited per kernel).
// Layout: depth / AO / history-length planes are fp32, normals float4 (xyz used); see DESIGN.md.
#include "dfx_common.cuh"
#include "dfx_pyramid.cuh"

namespace dfx
{

struct SsaoCam
{
    CamS c;
    Mat4 view;
};
__device__ __forceinline__ void stage_cam(SsaoCam& s, const dfx_camera_attribs* cams)
{
    if (threadIdx.x == 0 && threadIdx.y == 0)
    {
        load_cam(s.c, &cams[0]);
        load_mat(s.view, cams[0].mView);
    }
    __syncthreads();
}

// level-m row range of a full-resolution strip [y0, y1)
__host__ __device__ inline int mip_row(int y, int m, int full_h, int mip_h) { return y >= full_h ? mip_h : (y >> m); }

// A2 (prefiltered depth pyramid, SSAO_ComputePrefilteredDepthBuffer.fx:79-122): PrefilterOp in dfx_pyramid.cuh.

// ---------------------------------------------------------------------------------------------------------------------
// A3: ambient occlusion (SSAO_ComputeAmbientOcclusion.fx:132-231). One thread per pixel; the 18 depth taps per pixel are
// point-sampled from the prefiltered pyramid at mip = round(clamp(log2(|offset_px|) - DepthMIPSamplingOffset, 0, 4)).
// Background pixels keep the clear value 1.0 (the clear + discard of the reference are folded into the store).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_acos(float v) // :47-53
{
    float a = fabsf(v);
    float r = (-0.156583f * a + kHalfPi) * fsqrt(1.0f - a);
    return v >= 0.0f ? r : kPi - r;
}

__device__ __forceinline__ uint32_t occluded_sectors(float minH, float maxH, uint32_t bits) // :77-99
{
    minH = saturate(minH), maxH = saturate(maxH);
    if (maxH > minH)
    {
        uint32_t start = min((uint32_t)(minH * 32.0f), 31u);
        uint32_t end   = min((uint32_t)ceilf(maxH * 32.0f), 32u);
        if (end > start)
        {
            uint32_t n    = end - start;
            uint32_t mask = n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u);
            bits |= mask << start;
        }
    }
    return bits;
}

// one prefiltered-depth level, laid out for two 16-byte shared-memory loads
struct __align__(16) AoLevel
{
    const float* p;
    int          pitch, w;
    float        fw, fh;
    int          h, pad0;
};

template <int ALGO, bool N16>
__global__ void __launch_bounds__(256, DFX_OCC_AO) ssao_ao_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssao_attribs A, PyrView pyr,
                                                      TexRGBA<N16> normal, View<const float2> noise, View<float> out, int y0, int y1, int rev, int half, float self_offset)
{
    __shared__ SsaoCam S;
    __shared__ AoLevel lvl[DFX_MAX_MIPS];
    if (threadIdx.y == 1 && threadIdx.x < DFX_MAX_MIPS)
    {
        const int i = min((int)threadIdx.x, pyr.levels - 1);
        lvl[threadIdx.x] = AoLevel{pyr.lv[i].p, pyr.lv[i].pitch, pyr.lv[i].w, float(pyr.lv[i].w), float(pyr.lv[i].h), pyr.lv[i].h, 0};
    }
    stage_cam(S, cams);
    const CamS& cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;

    // The kernel is issue-bound (18 taps x view-space reconstruction per pixel), not bandwidth-bound, so the arithmetic is
    // written for instruction count: MUFU reciprocals / rsqrt / sincos, no precise-division or libm slow paths, and the
    // mip LOD log2() replaced by four squared-length threshold compares (floor(clamp(log2(l) - o, 0, 4) + 0.5) counts the
    // k in 1..4 with l >= 2^(o + k - 0.5)).
    // FEATURE_FLAG_HALF_RESOLUTION: the target (and the pyramid) is W/2 x H/2 and GetInvViewportSize() doubles (:68-75)
    const float ivs = half ? 2.0f : 1.0f;
    const float u = (float(x) + 0.5f) * (ivs * cam.ivw), v = (float(y) + 0.5f) * (ivs * cam.ivh);
    const float depth = __ldg(&pyr.lv[0].at(x, y)); // point sample at the pixel centre == Load(x, y)
    if (is_background(depth, rev))
    {
        st_cs(&out.at(x, y), 1.0f);
        return;
    }
    const float kx = 2.0f * frcp(cam.m00), ky = -2.0f * frcp(cam.m11); // view.xy = z * (uv - 0.5) * (kx, ky)
    auto to_view = [&](float su, float sv, float d) {
        const float z = fdiv(cam.m32 - d * cam.m33, d * cam.m23 - cam.m22);
        return make_float3(z * (su - 0.5f) * kx, z * (sv - 0.5f) * ky, z);
    };
    const float3 nvs = mul_dir(xyz(half ? sample_point_clamp(normal, u, v) : normal.ld(x, y)), S.view); // LoadNormalWS: point clamp
    float3       pvs = to_view(u, v, depth);
    pvs              = pvs + nvs * (self_offset * pvs.z); // 0.00001, or 0.005 with SSAO_OPTION_HALF_PRECISION_DEPTH (:145-150)
    const float3 view = -fnormalize(pvs);
    const float2 xi   = __ldg(&noise.at(x & 127, y & 127));
    const float  cxk = (u - 0.5f) * kx, cyk = (v - 0.5f) * ky;

    const float effectRadius = A.EffectRadius * A.RadiusMultiplier;
    const float falloffRange = A.EffectFalloffRange * effectRadius;
    const float falloffFrom  = effectRadius - falloffRange;
    const float falloffMul   = -1.0f / falloffRange;
    const float falloffAdd   = falloffFrom / falloffRange + 1.0f;
    float       sampleRadius = 0.5f * effectRadius * cam.m00;
    if (cam.m33 == 0.0f) sampleRadius = fdiv(sampleRadius, pvs.z); // perspective
    // squared pixel-length thresholds of mips 1..4
    float thr2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const float t = exp2f(A.DepthMIPSamplingOffset + float(k + 1) - 0.5f);
        thr2[k]       = t * t;
    }
    const int maxMip = pyr.levels - 1;

    float visibility = 0.0f;
#pragma unroll
    for (int slice = 0; slice < 3; ++slice)
    {
        const float phi = (xi.x + float(slice) / 3.0f) * kPi;
        float       so, co;
        __sincosf(phi, &so, &co);
        // sliceDir = (co, so, 0)
        const float  sdv        = co * view.x + so * view.y;
        const float3 orthoSlice = make_float3(co - sdv * view.x, so - sdv * view.y, -sdv * view.z);
        const float3 axis       = fnormalize(make_float3(so * view.z, -co * view.z, co * view.y - so * view.x)); // cross(sliceDir, view)
        const float3 projN      = nvs - axis * dot(nvs, axis);
        const float  projN2     = dot(projN, projN);
        const float  invLen     = frsqrt(projN2);
        const float  projNLen   = projN2 * invLen;
        const float  cosNorm    = saturate(dot(projN, view) * invLen);
        const float  N          = signf(dot(orthoSlice, projN)) * fast_acos(cosNorm);
        float        sinN, cosN;
        __sincosf(N, &sinN, &cosN);

        uint32_t bits = 0u;
        // cos(N + pi/2) = -sin N ; cos(N - pi/2) = sin N
        const float minCos0 = -sinN, minCos1 = sinN;
        float       maxCos0 = minCos0, maxCos1 = minCos1;

        float sdx = co * 0.5f * sampleRadius, sdy = so * -0.5f * sampleRadius;
        sdx *= cam.vh * cam.ivw; // aspect-ratio correction
        const float sdkx = sdx * kx, sdky = sdy * ky;

#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
            const float noiseS = fracf(xi.y + float(slice + s * 3) * 0.6180339887498948482f);
            const float smp    = (float(s) + noiseS) * (1.0f / 3.0f);
            const float offx = smp * smp * sdx, offy = smp * smp * sdy;
            const float lx = offx * cam.vw, ly = offy * cam.vh, l2 = lx * lx + ly * ly;
            int         mip = (l2 >= thr2[0]) + (l2 >= thr2[1]) + (l2 >= thr2[2]) + (l2 >= thr2[3]);
            mip             = min(mip, maxMip);
            const AoLevel lv = lvl[mip]; // two 16-byte LDS: pointer, pitch, size as int and as float
            const float u0 = u + offx, v0 = v + offy, u1 = u - offx, v1 = v - offy;
            const int   ax = min(max(__float2int_rd(u0 * lv.fw), 0), lv.w - 1), ay = min(max(__float2int_rd(v0 * lv.fh), 0), lv.h - 1);
            const int   bx = min(max(__float2int_rd(u1 * lv.fw), 0), lv.w - 1), by = min(max(__float2int_rd(v1 * lv.fh), 0), lv.h - 1);
            const float da = __ldg(lv.p + (unsigned)(ay * lv.pitch + ax)), db = __ldg(lv.p + (unsigned)(by * lv.pitch + bx));
            // view.xy = z * ((uv - 0.5) * k) with (uv - 0.5) * k = (centre - 0.5) * k +- smp^2 * (slice direction * k)
            const float za = fdiv(cam.m32 - da * cam.m33, da * cam.m23 - cam.m22), zb = fdiv(cam.m32 - db * cam.m33, db * cam.m23 - cam.m22);
            const float ox = smp * smp * sdkx, oy = smp * smp * sdky;
            const float3 d0 = make_float3(za * (cxk + ox), za * (cyk + oy), za) - pvs;
            const float3 d1 = make_float3(zb * (cxk - ox), zb * (cyk - oy), zb) - pvs;
            const float  q0 = dot(d0, d0), q1 = dot(d1, d1);
            const float  r0 = frsqrt(q0), r1 = frsqrt(q1);
            const float  l0 = q0 * r0, l1 = q1 * r1; // lengths
            const float  w0 = saturate(l0 * falloffMul + falloffAdd), w1 = saturate(l1 * falloffMul + falloffAdd);
            if (ALGO == DFX_SSAO_ALGORITHM_VBAO)
            {
                // ComputeSampleOcclusion :101-119
                const float3 thick = view * A.BitmaskThickness;
                float f0 = fast_acos(dot(d0, view) * r0), b0 = fast_acos(dot(fnormalize(d0 - thick), view));
                float f1 = fast_acos(dot(d1, view) * r1), b1 = fast_acos(dot(fnormalize(d1 - thick), view));
                const float nb = -N, ipi = 1.0f / kPi;
                f0 = saturate((-f0 - nb + kHalfPi) * ipi), b0 = saturate((-b0 - nb + kHalfPi) * ipi);
                f1 = saturate((f1 - nb + kHalfPi) * ipi), b1 = saturate((b1 - nb + kHalfPi) * ipi);
                if (w0 > 0.0f) bits = occluded_sectors(b0, f0, bits);
                if (w1 > 0.0f) bits = occluded_sectors(f1, b1, bits);
            }
            else
            {
                // ComputeSampleHorizons :121-130
                const float c0 = dot(d0, view) * r0, c1 = dot(d1, view) * r1;
                maxCos0 = fmaxf(maxCos0, lerpf(minCos0, c0, w0));
                maxCos1 = fmaxf(maxCos1, lerpf(minCos1, c1, w1));
            }
        }

        if (ALGO == DFX_SSAO_ALGORITHM_VBAO)
        {
            visibility += 1.0f - float(__popc(bits)) * (1.0f / 32.0f);
        }
        else
        {
            const float h0 = fast_acos(maxCos0), h1 = -fast_acos(maxCos1);
            if (ALGO == DFX_SSAO_ALGORITHM_HBAO)
                visibility += 0.5f * ((1.0f - __cosf(h0)) + (1.0f - __cosf(h1))); // IntegrateArcUniform :55-58
            else
            {
                // IntegrateArcCosWeighted :60-66
                const float H1 = h0 * 2.0f, H2 = h1 * 2.0f;
                visibility += projNLen * 0.25f * ((-__cosf(H1 - N) + cosNorm + H1 * sinN) + (-__cosf(H2 - N) + cosNorm + H2 * sinN));
            }
        }
    }
    st_cs(&out.at(x, y), visibility * (1.0f / 3.0f));
}

// ---------------------------------------------------------------------------------------------------------------------
// A0 (half resolution only): checkerboard of the 2x2 min / max depth — SSAO_ComputeDownsampledDepth.fx:8-29
// ---------------------------------------------------------------------------------------------------------------------
// VEC: the depth rows are 8-byte aligned (even pitch), one 64-bit load per row; otherwise (a caller's tightly pitched plane of odd
// width) four 32-bit loads.
template <bool VEC>
__global__ void __launch_bounds__(256) ssao_downsample_depth_kernel(View<const float> depth, View<float> out, int y0, int y1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out.w || y >= y1) return;
    float2 r0, r1;
    if (VEC)
    {
        r0 = __ldg(reinterpret_cast<const float2*>(depth.row(2 * y)) + x); // (2x, 2y), (2x+1, 2y)
        r1 = __ldg(reinterpret_cast<const float2*>(depth.row(2 * y + 1)) + x);
    }
    else
    {
        r0 = make_float2(__ldg(&depth.at(2 * x, 2 * y)), __ldg(&depth.at(2 * x + 1, 2 * y)));
        r1 = make_float2(__ldg(&depth.at(2 * x, 2 * y + 1)), __ldg(&depth.at(2 * x + 1, 2 * y + 1)));
    }
    const float  mn = fminf(fminf(r0.x, r1.x), fminf(r0.y, r1.y)), mx = fmaxf(fmaxf(r0.x, r1.x), fmaxf(r0.y, r1.y));
    out.at(x, y)    = lerpf(mn, mx, float((x + y) & 1));
}

// ---------------------------------------------------------------------------------------------------------------------
// A4 (half resolution only): 3x3 joint-bilateral upsampling of the half-res occlusion — SSAO_ComputeBilateralUpsampling.fx:62-139
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ssao_upsample_kernel(const dfx_camera_attribs* __restrict__ cams, View<const float> depth,
                                                            View<const float> occ, View<float> out, int y0, int y1, int rev)
{
    __shared__ SsaoCam S;
    stage_cam(S, cams);
    const CamS&   cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float dc = __ldg(&depth.at(x, y));
    if (is_background(dc, rev))
    {
        st_cs(&out.at(x, y), 1.0f);
        return;
    }
    const int   cx = x >> 1, cy = y >> 1;                       // int2(0.5 * floor(Position))
    const int   hw = (int)(0.5f * cam.vw), hh = (int)(0.5f * cam.vh);
    // IEEE division here: the depth weight is exp(-alpha^2 / 1.1e-4), a relative camera-Z error of 1e-6 is visible in it
    auto        depth_to_camz_precise = [&](float d, const CamS& c) { return (c.m32 - d * c.m33) / (d * c.m23 - c.m22); };
    const float zc = depth_to_camz_precise(dc, cam), izc = 1.0f / fmaxf(zc, 1e-6f);
    // ComputeSpatialWeight(d2, 0.9) = exp(-d2 / 1.62) for d2 = 0, 1, 2; depth weight exp(-alpha^2 / (2 * 0.0075^2))
    const float ws[3] = {1.0f, 0.53940751f, 0.29096046f};
    float       sum = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
        {
            const int   lx = min(max(cx + dx, 0), hw - 1), ly = min(max(cy + dy, 0), hh - 1);
            const float tu = 2.0f * (float(lx) + 0.5f) * cam.ivw, tv = 2.0f * (float(ly) + 0.5f) * cam.ivh;
            const float sig = load0(occ, lx, ly);
            const float gd  = sample_linear_clamp(depth, tu, tv);
            const float zg  = depth_to_camz_precise(gd, cam);
            const float a   = fabsf(zc - zg) * izc;
            const float wz  = expf(-(a * a) / (2.0f * 0.0075f * 0.0075f));
            const float w   = ws[dx * dx + dy * dy] * wz;
            sum += w * sig;
            wsum += w;
        }
    float r;
    if (wsum > 0.0f)
        r = sum / wsum;
    else
        r = sample_linear_clamp(occ, 2.0f * (float(cx) + 0.5f) * cam.ivw, 2.0f * (float(cy) + 0.5f) * cam.ivh);
    st_cs(&out.at(x, y), r);
}

// ---------------------------------------------------------------------------------------------------------------------
// A5: temporal accumulation (SSAO_ComputeTemporalAccumulation.fx:151-182). Background keeps the clear value (1, 1).
// ---------------------------------------------------------------------------------------------------------------------
struct TemporalCam
{
    CamS c, p;
};
__global__ void __launch_bounds__(256) ssao_temporal_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssao_attribs A,
                                                            View<const float> curr_occ, View<const float> prev_occ,
                                                            View<const float> prev_hist, View<const float> curr_depth,
                                                            View<const float> prev_depth, View<const float2> motion, View<float> out_occ,
                                                            View<float> out_hist, int y0, int y1, int rev)
{
    __shared__ TemporalCam S;
    if (threadIdx.x == 0 && threadIdx.y == 0) load_cam(S.c, &cams[0]), load_cam(S.p, &cams[1]);
    __syncthreads();
    const CamS& cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_occ.w || y >= y1) return;

    const float depth = __ldg(&curr_depth.at(x, y));
    if (is_background(depth, rev))
    {
        st_cs(&out_occ.at(x, y), 1.0f);
        st_cs(&out_hist.at(x, y), 1.0f);
        return;
    }
    const int    W = (int)cam.vw, H = (int)cam.vh;
    float2       mv = __ldg(&motion.at(x, y));
    mv.x *= 0.5f, mv.y *= -0.5f; // F3NDC_XYZ_TO_UVD_SCALE.xy
    const float plx = (float(x) + 0.5f) - mv.x * cam.vw, ply = (float(y) + 0.5f) - mv.y * cam.vh;

    // ComputeReprojection :105-149
    const float currZ = depth_to_camz(depth, cam);
    const Bilin b     = bilinear_uc(plx, ply, W, H);
    auto  similar = [&](int sx, int sy) {
        float pz = depth_to_camz(load0(prev_depth, sx, sy), S.p);
        return fabsf(1.0f - currZ / pz) < 0.01f ? 1.0f : 0.0f;
    };
    const float w00 = b.w00 * similar(b.x0, b.y0), w10 = b.w10 * similar(b.x1, b.y0);
    const float w01 = b.w01 * similar(b.x0, b.y1), w11 = b.w11 * similar(b.x1, b.y1);
    const float total = w00 * 1.0f + w10 * 1.0f + w01 * 1.0f + w11 * 1.0f;

    float rOcc = 1.0f, rHist = 1.0f;
    const bool ok = total > 0.01f && !A.ResetAccumulation;
    const float currOcc = __ldg(&curr_occ.at(x, y));
    if (ok)
    {
        const float o00 = load0(prev_occ, b.x0, b.y0), o10 = load0(prev_occ, b.x1, b.y0), o01 = load0(prev_occ, b.x0, b.y1), o11 = load0(prev_occ, b.x1, b.y1);
        const float h00 = fminf(load0(prev_hist, b.x0, b.y0) + 1.0f, 16.0f), h10 = fminf(load0(prev_hist, b.x1, b.y0) + 1.0f, 16.0f);
        const float h01 = fminf(load0(prev_hist, b.x0, b.y1) + 1.0f, 16.0f), h11 = fminf(load0(prev_hist, b.x1, b.y1) + 1.0f, 16.0f);
        rOcc  = (o00 * w00 + o10 * w10 + o01 * w01 + o11 * w11) / total;
        rHist = (h00 * w00 + h10 * w10 + h01 * w01 + h11 * w11) / total;

        // ComputePixelStatistic :81-103
        float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
            {
                float s = __ldg(&curr_occ.at(min(max(x + dx, 0), W - 1), min(max(y + dy, 0), H - 1)));
                m1 += s;
                m2 += s * s;
            }
        const float mean = m1 * (1.0f / 9.0f);
        const float var  = (m2 * (1.0f / 9.0f)) - (mean * mean);
        const float sd   = fsqrt(fmaxf(var, 0.0f));

        const float aspect = cam.vw * cam.ivh;
        const float mf     = saturate(1.025f - length(make_float2(mv.x * aspect, mv.y)) * 128.0f);
        const float gamma  = lerpf(0.5f, 2.5f, mf * mf);
        const float lo = mean - gamma * sd, hi = mean + gamma * sd;
        const bool  inside = lo < rOcc && rOcc < hi;
        rHist = inside ? rHist : fmaxf(1.0f, mf * rHist);
    }
    const float alpha = 1.0f / rHist;
    st_cs(&out_occ.at(x, y), lerpf(rOcc, currOcc, alpha));
    st_cs(&out_hist.at(x, y), rHist);
}

// A6 (convoluted AO-history and depth pyramids, SSAO_ComputeConvolutedDepthHistory.fx:93-109): ConvoluteOp in dfx_pyramid.cuh.

// ---------------------------------------------------------------------------------------------------------------------
// A7: resampled history (SSAO_ComputeResampledHistory.fx:56-115)
// ---------------------------------------------------------------------------------------------------------------------
DFX_HD float geometry_weight(float3 center, float3 tap, float3 n, float planeNorm) // SSAO_Common.fxh:25-28
{
    return saturate(1.0f - fabsf(dot(tap - center, n)) * planeNorm);
}

template <bool N16>
__global__ void __launch_bounds__(256) ssao_resample_kernel(const dfx_camera_attribs* __restrict__ cams, PyrView occ, PyrView dep,
                                                            View<const float> history, TexRGBA<N16> normal, View<float> out, int y0, int y1, int rev)
{
    __shared__ SsaoCam S;
    stage_cam(S, cams);
    const CamS& cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;

    // all three loads are issued before the decision: in steady state (history long enough for the early-out of :67-71) the pass is a
    // 16 B/px copy whose speed is set by how many loads are in flight per thread (ncu: issue-active 25 %, L1 hit 27 % with one at a time)
    const float depth = __ldg(&dep.lv[0].at(x, y));
    const float hist  = __ldg(&history.at(x, y));
    const float occ0  = __ldg(&occ.lv[0].at(x, y));
    const float acc   = (hist - 1.0f) / 4.0f;
    if (is_background(depth, rev) || acc >= 1.0f)
    {
        st_cs(&out.at(x, y), occ0);
        return;
    }
    int          mip = min((int)(4.0f * (1.0f - saturate(acc))), occ.levels - 1);
    const float  posx = float(x) + 0.5f, posy = float(y) + 0.5f;
    const float3 pvs  = screen_to_view(posx * cam.ivw, posy * cam.ivh, depth, cam);
    const float3 nvs  = mul_dir(xyz(normal.ld(x, y)), S.view);
    const float  planeNorm = 10.0f / (1.0f + depth_to_camz(depth, cam));

    float osum = 0.0f, wsum = 0.0f;
    while (mip >= 0 && wsum < 0.995f)
    {
        const float inv = 1.0f / float(1u << (unsigned)mip);
        const float rx = cam.vw * inv, ry = cam.vh * inv; // GetMipResolution: float, not the integer mip size
        const float lx = posx * inv, ly = posy * inv;
        const int   ix = (int)(lx - 0.5f), iy = (int)(ly - 0.5f);
        const float fx = fracf(lx + 0.5f), fy = fracf(ly + 0.5f);
        const float w[4] = {(1.0f - fx) * (1.0f - fy), fx * (1.0f - fy), (1.0f - fx) * fy, fx * fy};
        osum = 0.0f, wsum = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const int   tx = ix + (i & 1), ty = iy + (i >> 1);
            const float tu = (float(tx) + 0.5f) * (1.0f / rx), tv = (float(ty) + 0.5f) * (1.0f / ry);
            const float sd = sample_linear_clamp(dep.lv[mip], tu, tv);
            const float so = sample_point_clamp(occ.lv[mip], tu, tv);
            const float3 svs = screen_to_view(tu, tv, sd, cam);
            const float  wz  = geometry_weight(pvs, svs, nvs, planeNorm);
            osum += so * w[i] * wz;
            wsum += w[i] * wz;
        }
        --mip;
    }
    st_cs(&out.at(x, y), osum / wsum);
}

// ---------------------------------------------------------------------------------------------------------------------
// A8: spatial reconstruction, the "bilateral blur" (SSAO_ComputeSpatialReconstruction.fx:49-100): 8-tap Poisson disk
// rotated per pixel by Bayer4x4(frame), Gaussian x plane-distance weights, radius shrinking with accumulated history.
// ---------------------------------------------------------------------------------------------------------------------
__constant__ float3 kPoisson8[8] = {{-0.4706069f, -0.4427112f, +0.6461146f}, {-0.9057375f, +0.3003471f, +0.9542373f},
                                    {-0.3487388f, +0.4037880f, +0.5335386f}, {+0.1023042f, +0.6439373f, +0.6520134f},
                                    {+0.5699277f, +0.3513750f, +0.6695386f}, {+0.2939128f, -0.1131226f, +0.3149309f},
                                    {+0.7836658f, -0.4208784f, +0.8895339f}, {+0.1564120f, -0.8198990f, +0.8346850f}};

// exp(-z^2 / (2 * 0.9^2)) for the eight disk samples (the shader evaluates this constant expression per tap)
__constant__ float kPoisson8Weight[8] = {0.77283178f, 0.570022457f, 0.838854363f, 0.769187387f, 0.758268871f, 0.940613337f, 0.613583686f, 0.650469323f};

template <bool N16>
__global__ void __launch_bounds__(256) ssao_spatial_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssao_attribs A,
                                                           View<const float> occlusion, View<const float> history, View<const float> depth,
                                                           TexRGBA<N16> normal, View<float> out, int y0, int y1, int rev)
{
    __shared__ SsaoCam S;
    stage_cam(S, cams);
    const CamS& cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;

    const float hist = __ldg(&history.at(x, y));
    const float d    = __ldg(&depth.at(x, y));
    const float occC = __ldg(&occlusion.at(x, y));
    // acc = |(hist-1)/8|^0.2 >= 1  <=>  |hist-1| >= 8: decide the early-out without the pow
    const float hq = fabsf((hist - 1.0f) / 8.0f);
    if (is_background(d, rev) || hq >= 1.0f)
    {
        st_cs(&out.at(x, y), lerpf(1.0f, occC, A.AlphaInterpolation));
        return;
    }
    const float  acc = __powf(hq, 0.2f);
    const int    W = (int)cam.vw, H = (int)cam.vh;
    const float  posx = float(x) + 0.5f, posy = float(y) + 0.5f;
    const float  kx = 2.0f * frcp(cam.m00), ky = -2.0f * frcp(cam.m11);
    auto to_view = [&](float su, float sv, float dd) {
        const float z = fdiv(cam.m32 - dd * cam.m33, dd * cam.m23 - cam.m22);
        return make_float3(z * (su - 0.5f) * kx, z * (sv - 0.5f) * ky, z);
    };
    const float3 pvs  = to_view(posx * cam.ivw, posy * cam.ivh, d);
    const float3 nvs  = mul_dir(xyz(normal.ld(x, y)), S.view);
    float        rs, rc;
    __sincosf(2.0f * kPi * bayer4x4((uint32_t)x, (uint32_t)y, cam.frame_index), &rs, &rc);
    const float radius    = lerpf(0.0f, A.SpatialReconstructionRadius, 1.0f - saturate(acc));
    const float planeNorm = fdiv(10.0f, 1.0f + pvs.z);

    float osum = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const float3 P = kPoisson8[i];
        // RotateVector(Rotator=(cos, sin, -sin, cos), v) = v.x*(cos, -sin) + v.y*(sin, cos)
        const float xi = P.x * rc + P.y * rs, yi = P.x * -rs + P.y * rc;
        const int   sx = min(max((int)(posx + radius * xi), 0), W - 1), sy = min(max((int)(posy + radius * yi), 0), H - 1);
        const float sd = __ldg(&depth.at(sx, sy)), so = __ldg(&occlusion.at(sx, sy));
        const float3 svs = to_view((float(sx) + 0.5f) * cam.ivw, (float(sy) + 0.5f) * cam.ivh, sd);
        const float  ws  = kPoisson8Weight[i]; // exp(-z^2 / (2 * 0.9^2)), ComputeSpatialWeight with SSAO_SPATIAL_RECONSTRUCTION_SIGMA
        const float  wz  = geometry_weight(pvs, svs, nvs, planeNorm);
        osum += ws * wz * so;
        wsum += ws * wz;
    }
    const float o = wsum > 0.0f ? osum / wsum : occC;
    st_cs(&out.at(x, y), lerpf(1.0f, o, A.AlphaInterpolation));
}

// A8 with the tap window staged in shared memory. Every tap lands within +-4 pixels of its pixel (radius <= SpatialReconstructionRadius
// = 4, |Poisson sample| < 1), so a 32x8 CTA needs the 40x16 window of the depth and of the resampled AO: two TMA 2-D tile loads
// (cp.async.bulk.tensor; the parts of a box beyond the plane arrive as zeros and are never read - tap coordinates are clamped to
// the plane first). The depth window is converted to view-space Z once per texel instead of once per tap, and the plane distance
// dot(tapVS - centreVS, N) is evaluated as z_tap * dot(ray(tap), N) - dot(centreVS, N). CTAs in which no pixel needs the filter
// (background, or a history long enough for the early-out of :65-69 - the steady state) never issue the loads.
struct SpatialMaps
{
    CUtensorMap depth, occ;
};
constexpr int kSpTileW = 40, kSpTileH = 16;

template <bool N16>
__global__ void __launch_bounds__(256) ssao_spatial_tile_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssao_attribs A, const __grid_constant__ SpatialMaps maps,
                                                                View<const float> occlusion, View<const float> history, View<const float> depth,
                                                                TexRGBA<N16> normal, View<float> out, int y0, int y1, int rev)
{
    __shared__ SsaoCam                  S;
    __shared__ __align__(128) float     tz[kSpTileH][kSpTileW];
    __shared__ __align__(128) float     to[kSpTileH][kSpTileW];
    __shared__ __align__(8) uint64_t    bar;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid == 0) mbar_init(&bar, 1);
    stage_cam(S, cams); // __syncthreads inside
    const CamS&   cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    const bool    inside = x < out.w && y < y1;
    float         hist = 0.f, d = 0.f, occC = 0.f, hq = 2.f;
    bool          live = false;
    if (inside)
    {
        hist = __ldg(&history.at(x, y)), d = __ldg(&depth.at(x, y)), occC = __ldg(&occlusion.at(x, y));
        hq   = fabsf((hist - 1.0f) / 8.0f);
        live = !(is_background(d, rev) || hq >= 1.0f);
        if (!live) st_cs(&out.at(x, y), lerpf(1.0f, occC, A.AlphaInterpolation));
    }
    if (!__syncthreads_or(live)) return;
    const int X0 = int(blockIdx.x) * 32 - 4, Y0 = y0 + int(blockIdx.y) * 8 - 4;
    if (tid == 0)
    {
        mbar_arrive_expect_tx(&bar, uint32_t(sizeof(tz) + sizeof(to)));
        tma_load_2d(&tz[0][0], &maps.depth, X0, Y0, &bar);
        tma_load_2d(&to[0][0], &maps.occ, X0, Y0, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = tid; i < kSpTileW * kSpTileH; i += 256) // depth -> view-space Z, once per staged texel
    {
        float* p = &tz[0][0] + i;
        *p       = fdiv(cam.m32 - *p * cam.m33, *p * cam.m23 - cam.m22);
    }
    __syncthreads();
    if (!live) return;
    const float  acc = __powf(hq, 0.2f);
    const int    W = (int)cam.vw, H = (int)cam.vh;
    const float  posx = float(x) + 0.5f, posy = float(y) + 0.5f;
    const float  kx = 2.0f * frcp(cam.m00), ky = -2.0f * frcp(cam.m11);
    const float  zc = tz[y - Y0][x - X0];
    const float3 pvs = make_float3(zc * (posx * cam.ivw - 0.5f) * kx, zc * (posy * cam.ivh - 0.5f) * ky, zc);
    const float3 nvs = mul_dir(xyz(normal.ld(x, y)), S.view);
    const float  pn  = dot(pvs, nvs);
    float        rs, rc;
    __sincosf(2.0f * kPi * bayer4x4((uint32_t)x, (uint32_t)y, cam.frame_index), &rs, &rc);
    const float radius    = lerpf(0.0f, A.SpatialReconstructionRadius, 1.0f - saturate(acc));
    const float planeNorm = fdiv(10.0f, 1.0f + pvs.z);
    // ray(tap) . N = ((sx + 0.5) * ivw - 0.5) * kx * N.x + ((sy + 0.5) * ivh - 0.5) * ky * N.y + N.z, linear in the integer tap position
    const float gx = cam.ivw * kx * nvs.x, gy = cam.ivh * ky * nvs.y, g0 = (0.5f * cam.ivw - 0.5f) * kx * nvs.x + (0.5f * cam.ivh - 0.5f) * ky * nvs.y + nvs.z;
    float       osum = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const float3 P = kPoisson8[i];
        const float  xi = P.x * rc + P.y * rs, yi = P.x * -rs + P.y * rc;
        const int    sx = min(max((int)(posx + radius * xi), 0), W - 1), sy = min(max((int)(posy + radius * yi), 0), H - 1);
        const float  sz = tz[sy - Y0][sx - X0], so = to[sy - Y0][sx - X0];
        const float  dist = sz * (float(sx) * gx + float(sy) * gy + g0) - pn;
        const float  wz   = saturate(1.0f - fabsf(dist) * planeNorm);
        const float  ws   = kPoisson8Weight[i];
        osum += ws * wz * so;
        wsum += ws * wz;
    }
    const float o = wsum > 0.0f ? osum / wsum : occC;
    st_cs(&out.at(x, y), lerpf(1.0f, o, A.AlphaInterpolation));
}

static bool make_pyr(const dfx_pyramid* p, PyrView& v, int min_levels)
{
    if (!p || p->levels < min_levels || p->levels > DFX_MAX_MIPS) return false;
    v.levels = p->levels;
    for (int i = 0; i < p->levels; ++i)
    {
        if (!make_view<const float>(&p->level[i], DFX_FORMAT_R32F, v.lv[i])) return false;
        if (i > 0 && (v.lv[i].w != max(v.lv[0].w >> i, 1) || v.lv[i].h != max(v.lv[0].h >> i, 1))) return false;
    }
    return true;
}
static bool make_pyr_rw(const dfx_pyramid* p, PyrViewRW& v, int min_levels)
{
    if (!p || p->levels < min_levels || p->levels > DFX_MAX_MIPS) return false;
    v.levels = p->levels;
    for (int i = 0; i < p->levels; ++i)
    {
        if (!make_view<float>(&p->level[i], DFX_FORMAT_R32F, v.lv[i])) return false;
        if (i > 0 && (v.lv[i].w != max(v.lv[0].w >> i, 1) || v.lv[i].h != max(v.lv[0].h >> i, 1))) return false;
    }
    return true;
}
static View<const float> ro(const View<float>& v) { return View<const float>{v.p, v.pitch, v.w, v.h}; }

} // namespace dfx

using namespace dfx;

#define DFX_ROWS_ALIGNED(rows, h) DFX_REQUIRE(rows_ok(rows, h) && (rows.y0 % 64 == 0) && (rows.y1 % 64 == 0 || rows.y1 == (h)), "pyramid passes need 64-row aligned strips")

extern "C" dfx_status dfx_pass_ssao_prefilter_depth(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssao_attribs* attribs,
                                                    const dfx_pyramid* pyr, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_prefilter_depth");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    PyrViewRW P;
    DFX_REQUIRE(make_pyr_rw(pyr, P, 1), "bad prefiltered-depth pyramid");
    const int H = P.lv[0].h;
    DFX_ROWS_ALIGNED(rows, H);
    PyrPlanes<1> Q;
    Q.levels = P.levels;
    for (int i = 0; i < P.levels; ++i) Q.lv[0][i] = P.lv[i];
    return build_pyramid(stream, make_prefilter_op(cameras_dev, *attribs), Q, 4, rows, "ssao_prefilter pyramid kernel");
}

extern "C" dfx_status dfx_pass_ssao_ambient_occlusion(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssao_attribs* attribs,
                                                      const dfx_pyramid* prefiltered_depth, const dfx_plane* normal,
                                                      const dfx_plane* blue_noise_zw, const dfx_plane* occlusion, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_ambient_occlusion");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    PyrView P;
    DFX_REQUIRE(make_pyr(prefiltered_depth, P, 1), "bad prefiltered-depth pyramid");
    const int   rev = reversed_depth(&prefiltered_depth->level[0]); // level 0 is the depth buffer
    const float self_offset = (prefiltered_depth->level[0].flags & DFX_PLANE_FLAG_HALF_PRECISION_DEPTH) ? 0.005f : 0.00001f;
    DFX_TEX4(n, normal);
    DFX_VIEW(const float2, bn, blue_noise_zw, DFX_FORMAT_RG32F);
    DFX_VIEW(float, out, occlusion, DFX_FORMAT_R32F);
    // A pyramid of half the normal plane's size means FEATURE_FLAG_HALF_RESOLUTION (the reference allocates W/2 x H/2, …cpp:109-110)
    const int half = (P.lv[0].w != n.w || P.lv[0].h != n.h) ? 1 : 0;
    DFX_REQUIRE(!half || (P.lv[0].w == n.w / 2 && P.lv[0].h == n.h / 2), "the prefiltered pyramid must have the size of the normal plane or half of it");
    DFX_SAME_SIZE(P.lv[0], out);
    DFX_REQUIRE(bn.w == 128 && bn.h == 128, "blue noise must be 128x128");
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3 block(32, 8), grid(div_up(out.w, 32), div_up(rows.y1 - rows.y0, 8));
    switch (attribs->Algorithm)
    {
        case DFX_SSAO_ALGORITHM_GTAO: DFX_FMT16(is16(n), N16, ssao_ao_kernel<DFX_SSAO_ALGORITHM_GTAO, N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, P, n, bn, out, rows.y0, rows.y1, rev, half, self_offset)); break;
        case DFX_SSAO_ALGORITHM_HBAO: DFX_FMT16(is16(n), N16, ssao_ao_kernel<DFX_SSAO_ALGORITHM_HBAO, N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, P, n, bn, out, rows.y0, rows.y1, rev, half, self_offset)); break;
        case DFX_SSAO_ALGORITHM_VBAO: DFX_FMT16(is16(n), N16, ssao_ao_kernel<DFX_SSAO_ALGORITHM_VBAO, N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, P, n, bn, out, rows.y0, rows.y1, rev, half, self_offset)); break;
        default: return set_error(DFX_ERR_INVALID_ARG, "unknown SSAO algorithm %u", attribs->Algorithm);
    }
    DFX_LAUNCHED("ssao_ao_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssao_downsample_depth(void* stream, const dfx_plane* depth, const dfx_plane* out_half, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_downsample_depth");
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    DFX_VIEW(float, o, out_half, DFX_FORMAT_R32F);
    DFX_REQUIRE(o.w == d.w / 2 && o.h == d.h / 2, "the checkerboard plane must be width/2 x height/2 of the depth plane");
    DFX_REQUIRE(rows_ok(rows, o.h), "bad row range (rows of the half-resolution plane)");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3 block(32, 8), grid(div_up(o.w, 32), div_up(rows.y1 - rows.y0, 8));
    if (d.pitch % 2 == 0 && reinterpret_cast<uintptr_t>(d.p) % 8 == 0)
        ssao_downsample_depth_kernel<true><<<grid, block, 0, as_stream(stream)>>>(d, o, rows.y0, rows.y1);
    else
        ssao_downsample_depth_kernel<false><<<grid, block, 0, as_stream(stream)>>>(d, o, rows.y0, rows.y1);
    DFX_LAUNCHED("ssao_downsample_depth_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssao_upsample(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_plane* depth,
                                             const dfx_plane* occlusion_half, const dfx_plane* out_occlusion, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_upsample");
    DFX_REQUIRE(cameras_dev, "null argument");
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    DFX_VIEW(const float, oh, occlusion_half, DFX_FORMAT_R32F);
    DFX_VIEW(float, o, out_occlusion, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(d, o);
    DFX_REQUIRE(oh.w == d.w / 2 && oh.h == d.h / 2, "the half-resolution occlusion must be width/2 x height/2 of the depth plane");
    DFX_REQUIRE(rows_ok(rows, o.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3 block(32, 8), grid(div_up(o.w, 32), div_up(rows.y1 - rows.y0, 8));
    ssao_upsample_kernel<<<grid, block, 0, as_stream(stream)>>>(cameras_dev, d, oh, o, rows.y0, rows.y1, reversed_depth(depth));
    DFX_LAUNCHED("ssao_upsample_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssao_temporal(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssao_attribs* attribs,
                                             const dfx_plane* curr_occlusion, const dfx_plane* prev_occlusion,
                                             const dfx_plane* prev_history_length, const dfx_plane* reprojected_depth,
                                             const dfx_plane* previous_depth, const dfx_plane* closest_motion,
                                             const dfx_plane* out_occlusion, const dfx_plane* out_history_length, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_temporal");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const float, co, curr_occlusion, DFX_FORMAT_R32F);
    DFX_VIEW(const float, po, prev_occlusion, DFX_FORMAT_R32F);
    DFX_VIEW(const float, ph, prev_history_length, DFX_FORMAT_R32F);
    DFX_VIEW(const float, cd, reprojected_depth, DFX_FORMAT_R32F);
    const int rev = reversed_depth(reprojected_depth);
    DFX_VIEW(const float, pd, previous_depth, DFX_FORMAT_R32F);
    DFX_VIEW(const float2, mv, closest_motion, DFX_FORMAT_RG32F);
    DFX_VIEW(float, oo, out_occlusion, DFX_FORMAT_R32F);
    DFX_VIEW(float, oh, out_history_length, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(co, po);
    DFX_SAME_SIZE(co, ph);
    DFX_SAME_SIZE(co, cd);
    DFX_SAME_SIZE(co, pd);
    DFX_SAME_SIZE(co, mv);
    DFX_SAME_SIZE(co, oo);
    DFX_SAME_SIZE(co, oh);
    DFX_REQUIRE(rows_ok(rows, co.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3 block(32, 8), grid(div_up(co.w, 32), div_up(rows.y1 - rows.y0, 8));
    ssao_temporal_kernel<<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, co, po, ph, cd, pd, mv, oo, oh, rows.y0, rows.y1, rev);
    DFX_LAUNCHED("ssao_temporal_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssao_convolute(void* stream, const dfx_pyramid* occlusion_pyr, const dfx_pyramid* depth_pyr, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_convolute");
    PyrViewRW O, D;
    DFX_REQUIRE(make_pyr_rw(occlusion_pyr, O, 1) && make_pyr_rw(depth_pyr, D, 1), "bad pyramid");
    DFX_REQUIRE(O.levels == D.levels, "pyramid level mismatch");
    DFX_SAME_SIZE(O.lv[0], D.lv[0]);
    const int H = O.lv[0].h;
    DFX_ROWS_ALIGNED(rows, H);
    PyrPlanes<2> Q;
    Q.levels = O.levels;
    for (int i = 0; i < O.levels; ++i) Q.lv[0][i] = O.lv[i], Q.lv[1][i] = D.lv[i];
    return build_pyramid(stream, ConvoluteOp{}, Q, 4, rows, "ssao_convolute pyramid kernel");
}

extern "C" dfx_status dfx_pass_ssao_resample(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_pyramid* occlusion_pyr,
                                             const dfx_pyramid* depth_pyr, const dfx_plane* history_length, const dfx_plane* normal,
                                             const dfx_plane* out_occlusion, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_resample");
    DFX_REQUIRE(cameras_dev, "null argument");
    PyrView O, D;
    DFX_REQUIRE(make_pyr(occlusion_pyr, O, 1) && make_pyr(depth_pyr, D, 1), "bad pyramid");
    const int rev = reversed_depth(&depth_pyr->level[0]); // level 0 is the depth buffer
    DFX_REQUIRE(O.levels == D.levels, "pyramid level mismatch");
    DFX_VIEW(const float, h, history_length, DFX_FORMAT_R32F);
    DFX_TEX4(n, normal);
    DFX_VIEW(float, out, out_occlusion, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(O.lv[0], D.lv[0]);
    DFX_SAME_SIZE(O.lv[0], h);
    DFX_SAME_SIZE(O.lv[0], n);
    DFX_SAME_SIZE(O.lv[0], out);
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3 block(32, 8), grid(div_up(out.w, 32), div_up(rows.y1 - rows.y0, 8));
    DFX_FMT16(is16(n), N16, ssao_resample_kernel<N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, O, D, h, n, out, rows.y0, rows.y1, rev));
    DFX_LAUNCHED("ssao_resample_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssao_spatial(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssao_attribs* attribs,
                                            const dfx_plane* occlusion, const dfx_plane* history_length, const dfx_plane* depth,
                                            const dfx_plane* normal, const dfx_plane* out_occlusion, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssao_spatial");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const float, o, occlusion, DFX_FORMAT_R32F);
    DFX_VIEW(const float, h, history_length, DFX_FORMAT_R32F);
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    const int rev = reversed_depth(depth);
    DFX_TEX4(n, normal);
    DFX_VIEW(float, out, out_occlusion, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(o, h);
    DFX_SAME_SIZE(o, d);
    DFX_SAME_SIZE(o, n);
    DFX_SAME_SIZE(o, out);
    DFX_REQUIRE(rows_ok(rows, out.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3 block(32, 8), grid(div_up(out.w, 32), div_up(rows.y1 - rows.y0, 8));
    // dfx_tune("ssao_spatial_impl"): 1 (default) = tap window staged by TMA, 0 = taps gathered from HBM / L1
    const CUtensorMap *md = nullptr, *mo = nullptr;
    if (tune("ssao_spatial_impl", 1) == 1 && attribs->SpatialReconstructionRadius <= 4.0f)
    {
        md = tensor_map_r32f(View<float>{const_cast<float*>(d.p), d.pitch, d.w, d.h}, kSpTileW, kSpTileH);
        mo = tensor_map_r32f(View<float>{const_cast<float*>(o.p), o.pitch, o.w, o.h}, kSpTileW, kSpTileH);
    }
    if (md && mo)
        DFX_FMT16(is16(n), N16, ssao_spatial_tile_kernel<N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, SpatialMaps{*md, *mo}, o, h, d, n, out, rows.y0, rows.y1, rev));
    else
        DFX_FMT16(is16(n), N16, ssao_spatial_kernel<N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, o, h, d, n, out, rows.y0, rows.y1, rev));
    DFX_LAUNCHED("ssao_spatial_kernel");
    return DFX_OK;
}
