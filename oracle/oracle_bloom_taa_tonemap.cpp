// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h). Pinned against the reference's own shaders (oracle/refshader, tests/test_reference_shaders.py).
// Bloom B1-B4, TAA T0/T1, compose (reduced form) and ToneMap M1/M2 restated from the reference HLSL / C++.
#include "oracle.h"

namespace orc
{

// =====================================================================================================================
// Bloom.  Sampler for B1/B2: linear + border(0) (Bloom.cpp:52-59, :185, :219 — BorderSamplingModeSupported branch);
// B3/B4: linear clamp (Bloom.cpp:253-254). Pixel-centre UV of the OUTPUT target: NormalizedDeviceXYToTexUV(f2NormalizedXY)
// == (pixel + 0.5) / output size.
// =====================================================================================================================

static inline float3 tap(const TexF4& t, float2 uv, float2 off, Address a) { return sample_linear(t, uv + off, a).xyz(); }

int bloom_mip_count(int width, int height, float radius) // Bloom.cpp:152-156
{
    return int(radius * float(compute_mip_levels_count(width, height)));
}

struct Taps13
{
    float3 A, B, C, D, E, F, G, H, I, J, K, L, M;
};
static inline Taps13 taps13(const TexF4& in, float2 uv, Address addr)
{
    float2 ts(rcp(float(in.w)), rcp(float(in.h)));
    Taps13 t;
    t.A = tap(in, uv, ts * float2(-2.0f, +2.0f), addr);
    t.B = tap(in, uv, ts * float2(+0.0f, +2.0f), addr);
    t.C = tap(in, uv, ts * float2(+2.0f, +2.0f), addr);
    t.D = tap(in, uv, ts * float2(-2.0f, +0.0f), addr);
    t.E = tap(in, uv, ts * float2(+0.0f, +0.0f), addr);
    t.F = tap(in, uv, ts * float2(+2.0f, +0.0f), addr);
    t.G = tap(in, uv, ts * float2(-2.0f, -2.0f), addr);
    t.H = tap(in, uv, ts * float2(+0.0f, -2.0f), addr);
    t.I = tap(in, uv, ts * float2(+2.0f, -2.0f), addr);
    t.J = tap(in, uv, ts * float2(-1.0f, +1.0f), addr);
    t.K = tap(in, uv, ts * float2(+1.0f, +1.0f), addr);
    t.L = tap(in, uv, ts * float2(-1.0f, -1.0f), addr);
    t.M = tap(in, uv, ts * float2(+1.0f, -1.0f), addr);
    return t;
}

// Bloom_ComputePrefilteredTexture.fx:37-83
void bloom_prefilter(const dfx_bloom_attribs& A, const TexF4& color, TexF4& out, int threads)
{
    const int OW = std::max(color.w / 2, 1), OH = std::max(color.h / 2, 1);
    out.resize(OW, OH);
    parallel_rows(0, OH, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < OW; ++x)
            {
                float2 uv((float(x) + 0.5f) / float(OW), (float(y) + 0.5f) / float(OH));
                Taps13 t = taps13(color, uv, Address::Border);
                const float Weights[5] = {0.125f, 0.125f, 0.125f, 0.125f, 0.5f};
                float3      Groups[5];
                Groups[0] = (t.A + t.B + t.D + t.E) / 4.0f;
                Groups[1] = (t.B + t.C + t.E + t.F) / 4.0f;
                Groups[2] = (t.D + t.E + t.G + t.H) / 4.0f;
                Groups[3] = (t.E + t.F + t.H + t.I) / 4.0f;
                Groups[4] = (t.J + t.K + t.L + t.M) / 4.0f;
                float4 ColorSum;
                for (int g = 0; g < 5; ++g)
                {
                    float Weight = Weights[g] * (1.0f / (1.0f + Luminance(Groups[g]))); // KarisAverage :19-22
                    ColorSum += float4(Groups[g], 1.0f) * Weight;
                }
                float3 Color = ColorSum.xyz() / (ColorSum.w + 1.0e-5f);
                // Prefilter :24-35
                float Brightness = hmax(Color.x, hmax(Color.y, Color.z));
                float Knee       = A.Threshold * A.SoftTreshold;
                float Soft       = Brightness - A.Threshold + Knee;
                Soft             = clampf(Soft, 0.0f, 2.0f * Knee);
                Soft             = Soft * Soft * 0.25f / (Knee + 1.0e-5f);
                float Contribution = hmax(Soft, Brightness - A.Threshold);
                Contribution /= hmax(Brightness, 1.0e-5f);
                out.at(x, y) = float4(Color * Contribution, 0.0f);
            }
    });
}

// Bloom_ComputeDownsampledTexture.fx:11-41
void bloom_downsample(const TexF4& in, TexF4& out, int threads)
{
    const int OW = std::max(in.w / 2, 1), OH = std::max(in.h / 2, 1);
    out.resize(OW, OH);
    parallel_rows(0, OH, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < OW; ++x)
            {
                float2 uv((float(x) + 0.5f) / float(OW), (float(y) + 0.5f) / float(OH));
                Taps13 t = taps13(in, uv, Address::Border);
                float3 OutColor(0.0f, 0.0f, 0.0f);
                OutColor += (t.A + t.C + t.G + t.I) * 0.03125f;
                OutColor += (t.B + t.D + t.F + t.H) * 0.0625f;
                OutColor += (t.E + t.J + t.K + t.L + t.M) * 0.125f;
                out.at(x, y) = float4(OutColor, 0.0f);
            }
    });
}

static inline float3 tent9(const TexF4& lo, float2 uv)
{
    float2 ts(rcp(float(lo.w)), rcp(float(lo.h)));
    float3 A = tap(lo, uv, ts * float2(-1.0f, +1.0f), Address::Clamp);
    float3 B = tap(lo, uv, ts * float2(+0.0f, +1.0f), Address::Clamp);
    float3 C = tap(lo, uv, ts * float2(+1.0f, +1.0f), Address::Clamp);
    float3 D = tap(lo, uv, ts * float2(-1.0f, +0.0f), Address::Clamp);
    float3 E = tap(lo, uv, ts * float2(+0.0f, +0.0f), Address::Clamp);
    float3 F = tap(lo, uv, ts * float2(+1.0f, +0.0f), Address::Clamp);
    float3 G = tap(lo, uv, ts * float2(-1.0f, -1.0f), Address::Clamp);
    float3 H = tap(lo, uv, ts * float2(+0.0f, -1.0f), Address::Clamp);
    float3 I = tap(lo, uv, ts * float2(+1.0f, -1.0f), Address::Clamp);
    float3 ColorSum = E * 0.25f;
    ColorSum += (B + D + F + H) * 0.125f;
    ColorSum += (A + C + G + I) * 0.0625f;
    return ColorSum;
}

// Bloom_ComputeUpsampledTexture.fx:20-54, uInstID == 0
void bloom_upsample(const TexF4& same_level_down, const TexF4& coarser, TexF4& out, int threads)
{
    const int OW = same_level_down.w, OH = same_level_down.h;
    out.resize(OW, OH);
    parallel_rows(0, OH, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < OW; ++x)
            {
                float2 uv((float(x) + 0.5f) / float(OW), (float(y) + 0.5f) / float(OH));
                float3 ColorSum    = tent9(coarser, uv);
                float3 SourceColor = sample_linear(same_level_down, uv, Address::Clamp).xyz();
                out.at(x, y)       = float4(SourceColor + ColorSum, 0.0f);
            }
    });
}

// Bloom_ComputeUpsampledTexture.fx:45-48, uInstID != 0
void bloom_composite(const dfx_bloom_attribs& A, const TexF4& color, const TexF4& up0, TexF4& out, int threads)
{
    const int OW = color.w, OH = color.h;
    out.resize(OW, OH);
    parallel_rows(0, OH, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < OW; ++x)
            {
                float2 uv((float(x) + 0.5f) / float(OW), (float(y) + 0.5f) / float(OH));
                float3 ColorSum    = tent9(up0, uv);
                float3 SourceColor = sample_linear(color, uv, Address::Clamp).xyz();
                out.at(x, y) = float4(lerp(SourceColor, SourceColor + A.Intensity * ColorSum, A.AlphaInterpolation), 0.0f);
            }
    });
}

// =====================================================================================================================
// TAA
// =====================================================================================================================

float halton_sequence(uint Base, uint Index) // TemporalAntiAliasing.cpp:43-54
{
    float Result = 0.0f, F = 1.0f;
    while (Index > 0)
    {
        F      = F / float(Base);
        Result = Result + F * float(Index % Base);
        Index  = uint(std::floor(float(Index) / float(Base)));
    }
    return Result;
}

float2 taa_jitter_offset(uint frame_index, uint width, uint height) // TemporalAntiAliasing.cpp:63-78
{
    const uint SampleCount = 16u;
    float JitterX = (halton_sequence(2u, (frame_index % SampleCount) + 1) - 0.5f) / (0.5f * float(width));
    float JitterY = (halton_sequence(3u, (frame_index % SampleCount) + 1) - 0.5f) / (0.5f * float(height));
    return float2(JitterX, JitterY);
}

namespace
{
inline float3 RGBToYCoCg(float3 RGB, bool ycocg) // TAA_…fx:34-50
{
    if (!ycocg) return RGB;
    float Co = RGB.x - RGB.z, Temp = RGB.z + 0.5f * Co, Cg = RGB.y - Temp, Y = Temp + 0.5f * Cg;
    return float3(Y, Co, Cg);
}
inline float3 YCoCgToRGB(float3 c, bool ycocg) // :52-67
{
    if (!ycocg) return c;
    float Tmp = c.x - 0.5f * c.z, G = c.z + Tmp, B = Tmp - 0.5f * c.y, R = B + c.y;
    return float3(R, G, B);
}
inline float3 HDRToSDR(float3 c) { return c * float3(rcp(1.0f + c.x), rcp(1.0f + c.y), rcp(1.0f + c.z)); }                                        // :69-72
inline float3 SDRToHDR(float3 c) { return c * float3(rcp(1.0f - c.x + FLT_EPS_F), rcp(1.0f - c.y + FLT_EPS_F), rcp(1.0f - c.z + FLT_EPS_F)); } // :74-77

// :98-106 ClipToAABB ; Less/GreaterEqual -> 1.0/0.0 selectors fed to lerp (a + t*(b-a)); min with NaN returns the other operand
inline float3 ClipToAABB(float3 ColorPrev, float3 ColorCurr, float3 AABBCentre, float3 AABBExtents)
{
    const float MaxT = 10.0f;
    float3 Direction    = ColorCurr - ColorPrev;
    float3 Intersection = ((AABBCentre - sign3(Direction) * AABBExtents) - ColorPrev) / Direction;
    float3 ge(Intersection.x >= 0.0f ? 1.0f : 0.0f, Intersection.y >= 0.0f ? 1.0f : 0.0f, Intersection.z >= 0.0f ? 1.0f : 0.0f);
    float3 PossibleT = lerp(float3(MaxT + 1.0f, MaxT + 1.0f, MaxT + 1.0f), Intersection, ge);
    float  T  = hmin(MaxT, hmin(PossibleT.x, hmin(PossibleT.y, PossibleT.z)));
    float  lt = T < MaxT ? 1.0f : 0.0f;
    return lerp(ColorPrev, ColorPrev + Direction * T, float3(lt, lt, lt));
}
} // namespace

void taa_accumulate(const Camera& curr, const Camera& prev, const dfx_taa_attribs& A, uint flags, const TexF4& curr_color,
                    const TexF4& prev_accum, const TexF2& closest_motion, const TexF& reprojected_depth, const TexF& previous_depth,
                    TexF4& out, int threads)
{
    const int  W = curr_color.w, H = curr_color.h;
    const bool Gaussian = (flags & DFX_TAA_FEATURE_FLAG_GAUSSIAN_WEIGHTING) != 0;
    const bool Bicubic  = (flags & DFX_TAA_FEATURE_FLAG_BICUBIC_FILTER) != 0;
    const bool YCoCg    = (flags & DFX_TAA_FEATURE_FLAG_YCOCG_COLOR_SPACE) != 0;
    out.resize(W, H);
    const float2 Viewport(curr.f4ViewportSize.x, curr.f4ViewportSize.y);
    const float2 TexelSize(curr.f4ViewportSize.z, curr.f4ViewportSize.w);
    const int2   Dim(int(curr.f4ViewportSize.x), int(curr.f4ViewportSize.y));

    auto SampleCurrColor = [&](int x, int y) { return max3(curr_color.load(x, y).xyz(), 0.0f); };
    auto SamplePrev      = [&](float2 uv) { return sample_linear(prev_accum, uv, Address::Clamp); }; // Sam_LinearClamp (…cpp:234)

    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                float2 Motion       = closest_motion.load(px, py) * float2(F3NDC_XYZ_TO_UVD_SCALE.x, F3NDC_XYZ_TO_UVD_SCALE.y);
                float2 PrevPosition = Position - Motion * Viewport;

                if (!IsInsideScreen(PrevPosition, Viewport) || A.ResetAccumulation)
                {
                    out.at(px, py) = float4(SampleCurrColor(px, py), 0.5f);
                    continue;
                }

                float AspectRatio  = curr.f4ViewportSize.x * curr.f4ViewportSize.w;
                float MotionFactor = saturate(1.0f - length(float2(Motion.x * AspectRatio, Motion.y)) * 256.0f);

                // ComputeDepthDisocclusion :117-136 (unclamped Loads)
                float DepthFactor;
                {
                    int2  PrevPositioni(ftoi(PrevPosition.x), ftoi(PrevPosition.y));
                    float CurrDepth    = reprojected_depth.load(px, py);
                    float Disocclusion = 0.0f;
                    for (int y = -1; y <= 1; y++)
                        for (int x = -1; x <= 1; x++)
                        {
                            float PrevDepth = previous_depth.load(PrevPositioni.x + x, PrevPositioni.y + y);
                            // ComputeDepthDisocclusionWeight :108-115
                            float LinearDepthCurr  = std::fabs(DepthToCameraZ(CurrDepth, curr.mProj));
                            float LinearDepthPrev  = std::fabs(DepthToCameraZ(PrevDepth, prev.mProj));
                            float MaxLinearDepth   = hmax(LinearDepthCurr, LinearDepthPrev);
                            float LinearDepthDelta = std::fabs(LinearDepthCurr - LinearDepthPrev);
                            float Weight = std::exp(-LinearDepthDelta / hmax(MaxLinearDepth, 1e-6f));
                            Disocclusion = hmax(Disocclusion, Weight);
                        }
                    DepthFactor = Disocclusion > 0.9f ? 1.0f : 0.0f;
                }

                float3 RGBHDRCurrColor = SampleCurrColor(px, py);
                float4 RGBHDRPrevColor;
                if (Bicubic)
                {
                    // SamplePrevColorCatmullRom :138-173
                    float2 CenterPosition = floor2(PrevPosition - float2(0.5f, 0.5f)) + float2(0.5f, 0.5f);
                    float2 F  = PrevPosition - CenterPosition;
                    float2 F2 = F * F;
                    float2 F3 = F2 * F;
                    float2 W0 = -0.5f * F3 + F2 - 0.5f * F;
                    float2 W1 = 1.5f * F3 - 2.5f * F2 + float2(1.0f, 1.0f);
                    float2 W2 = -1.5f * F3 + 2.0f * F2 + 0.5f * F;
                    float2 W3 = 0.5f * F3 - 0.5f * F2;
                    float2 W12 = W1 + W2;
                    float2 TexPos0  = (CenterPosition - float2(1.0f, 1.0f)) * TexelSize;
                    float2 TexPos3  = (CenterPosition + float2(2.0f, 2.0f)) * TexelSize;
                    float2 TexPos12 = (CenterPosition + W2 / W12) * TexelSize;
                    float  P0 = W12.x * W0.y, P1 = W0.x * W12.y, P2 = W12.x * W12.y, P3 = W3.x * W12.y, P4 = W12.x * W3.y;
                    float4 Result;
                    Result += SamplePrev(float2(TexPos12.x, TexPos0.y)) * P0;
                    Result += SamplePrev(float2(TexPos0.x, TexPos12.y)) * P1;
                    Result += SamplePrev(float2(TexPos12.x, TexPos12.y)) * P2;
                    Result += SamplePrev(float2(TexPos3.x, TexPos12.y)) * P3;
                    Result += SamplePrev(float2(TexPos12.x, TexPos3.y)) * P4;
                    RGBHDRPrevColor = max4(Result * rcp(P0 + P1 + P2 + P3 + P4), 0.0f);
                }
                else
                {
                    RGBHDRPrevColor = max4(SamplePrev(PrevPosition * TexelSize), 0.0f); // :175-178
                }

                float3 YCoCgSDRCurrColor = RGBToYCoCg(HDRToSDR(RGBHDRCurrColor), YCoCg);
                float3 YCoCgSDRPrevColor = RGBToYCoCg(HDRToSDR(RGBHDRPrevColor.xyz()), YCoCg);

                auto ComputeCorrectedAlpha = [&](float Alpha) { return hmin(A.TemporalStabilityFactor, saturate(1.0f / (2.0f - Alpha))); };

                if (A.SkipRejection)
                {
                    float3 o = SDRToHDR(YCoCgToRGB(lerp(YCoCgSDRCurrColor, YCoCgSDRPrevColor, RGBHDRPrevColor.w), YCoCg));
                    out.at(px, py) = float4(o, ComputeCorrectedAlpha(RGBHDRPrevColor.w));
                    continue;
                }

                float VarianceGamma = lerp(0.75f, 2.5f, MotionFactor * MotionFactor);
                // ComputePixelStatisticYCoCgSDR :191-222
                float3 Mean, StdDev;
                {
                    float  WeightSum = 0.0f;
                    float3 M1, M2;
                    for (int x = -1; x <= 1; x++)
                        for (int y = -1; y <= 1; y++)
                        {
                            int2   L   = ClampScreenCoord(int2(px + x, py + y), Dim);
                            float3 SDR = RGBToYCoCg(HDRToSDR(SampleCurrColor(L.x, L.y)), YCoCg);
                            float  Weight = Gaussian ? std::exp(-3.0f * float(x * x + y * y) / ((1.0f + 1.0f) * (1.0f + 1.0f))) : 1.0f;
                            M1 += SDR * Weight;
                            M2 += SDR * SDR * Weight;
                            WeightSum += Weight;
                        }
                    Mean            = M1 / WeightSum;
                    float3 Variance = M2 / WeightSum - (Mean * Mean);
                    StdDev          = sqrt3(max3(Variance, 0.0f));
                }
                float3 Clamped = ClipToAABB(YCoCgSDRPrevColor, YCoCgSDRCurrColor, Mean, VarianceGamma * StdDev);
                float  Alpha   = RGBHDRPrevColor.w * MotionFactor * DepthFactor;
                float3 o       = SDRToHDR(YCoCgToRGB(lerp(YCoCgSDRCurrColor, Clamped, Alpha), YCoCg));
                out.at(px, py) = float4(o, ComputeCorrectedAlpha(Alpha));
            }
    });
}

// =====================================================================================================================
// compose — reduced form of Hydrogent/shaders/HnPostProcess.psh:145-185:
//   Color.rgb += SSR.rgb * SSR.a * SSRScale ;  Color.rgb *= lerp(1, AO, SSAOScale) ; alpha passes through.
// =====================================================================================================================
void compose(const TexF4& color, const TexF4* ssr, const TexF* ao, float ssr_scale, float ssao_scale, TexF4& out, int threads)
{
    out.resize(color.w, color.h);
    parallel_rows(0, color.h, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < color.w; ++x)
            {
                float4 C = color.load(x, y);
                float3 c = C.xyz();
                if (ssr && ssr_scale > 0.0f)
                {
                    float4 S = ssr->load(x, y);
                    c        = c + S.xyz() * S.w * ssr_scale;
                }
                if (ao && ssao_scale > 0.0f)
                {
                    float Occlusion = lerp(1.0f, ao->load(x, y), ssao_scale);
                    c               = c * Occlusion;
                }
                out.at(x, y) = float4(c, C.w);
            }
    });
}

// =====================================================================================================================
// ToneMapping.fxh
// =====================================================================================================================
float3 uncharted2_tonemap(float3 x) // :8-19
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + float3(C * B, C * B, C * B)) + float3(D * E, D * E, D * E)) / (x * (A * x + float3(B, B, B)) + float3(D * F, D * F, D * F))) -
           float3(E / F, E / F, E / F);
}

namespace
{
const float3 RGB_TO_LUMINANCE(0.212671f, 0.715160f, 0.072169f);

inline float3 mul33(const float r[3][3], float3 v) // mul(M, v), M built with MatrixFromRows
{
    return float3(r[0][0] * v.x + r[0][1] * v.y + r[0][2] * v.z, r[1][0] * v.x + r[1][1] * v.y + r[1][2] * v.z,
                  r[2][0] * v.x + r[2][1] * v.y + r[2][2] * v.z);
}
inline float3 AgXDefaultContrastApprox(float3 x) // :21-34
{
    float3 x2 = x * x, x4 = x2 * x2;
    return 15.5f * x4 * x2 - 40.14f * x4 * x + 31.96f * x4 - 6.868f * x2 * x + 0.4298f * x2 + 0.1191f * x - float3(0.00232f, 0.00232f, 0.00232f);
}
inline float3 AgX(float3 Color) // :36-58
{
    static const float M[3][3] = {{0.842479062253094f, 0.0784335999999992f, 0.0792237451477643f},
                                  {0.0423282422610123f, 0.878468636469772f, 0.0791661274605434f},
                                  {0.0423756549057051f, 0.0784336f, 0.879142973793104f}};
    const float MinEv = -12.47393f, MaxEv = 4.026069f;
    Color = mul33(M, Color);
    Color = float3(clampf(std::log2(Color.x), MinEv, MaxEv), clampf(std::log2(Color.y), MinEv, MaxEv), clampf(std::log2(Color.z), MinEv, MaxEv));
    Color = (Color - float3(MinEv, MinEv, MinEv)) / (MaxEv - MinEv);
    return AgXDefaultContrastApprox(Color);
}
inline float3 AgXEotf(float3 Color) // :60-74
{
    static const float M[3][3] = {{+1.19687900512017f, -0.0980208811401368f, -0.0990297440797205f},
                                  {-0.0528968517574562f, +1.15190312990417f, -0.0989611768448433f},
                                  {-0.0529716355144438f, -0.0980434501171241f, +1.15107367264116f}};
    Color = mul33(M, Color);
    return SRGBToLinear(Color);
}
inline float3 AgXPunchyLook(float3 Color, float fSaturation, float fOffset, float fSlope, float fPower) // :76-88
{
    float  Lum = dot(Color, RGB_TO_LUMINANCE);
    float3 c   = Color * fSlope + float3(fOffset, fOffset, fOffset);
    c          = float3(std::pow(c.x, fPower), std::pow(c.y, fPower), std::pow(c.z, fPower));
    return float3(Lum, Lum, Lum) + fSaturation * (c - float3(Lum, Lum, Lum));
}
} // namespace

float3 tone_map(float3 f3Color, const dfx_tonemap_attribs& At, float fAveLogLum) // :87-226
{
    float middleGray = At.fMiddleGray;
    float fLumScale  = middleGray / fAveLogLum;
    f3Color          = max3(f3Color, 0.0f);
    float  fInitialPixelLum = hmax(dot(RGB_TO_LUMINANCE, f3Color), 1e-10f);
    float  fScaledPixelLum  = fInitialPixelLum * fLumScale;
    float3 f3ScaledColor    = f3Color * fLumScale;
    float  whitePoint       = At.fWhitePoint;
    auto   satpow           = [&](float3 c) { return pow3(c, At.fLuminanceSaturation); };

    switch (At.iToneMappingMode)
    {
        case DFX_TONE_MAPPING_MODE_EXP:
        {
            float fToneMappedLum = 1.0f - std::exp(-fScaledPixelLum);
            return fToneMappedLum * satpow(f3Color / fInitialPixelLum);
        }
        case DFX_TONE_MAPPING_MODE_REINHARD:
        case DFX_TONE_MAPPING_MODE_REINHARD_MOD:
        {
            float L_xy = fScaledPixelLum;
            float fToneMappedLum = At.iToneMappingMode == DFX_TONE_MAPPING_MODE_REINHARD ? L_xy / (1.0f + L_xy) :
                                                                                           L_xy * (1.0f + L_xy / (whitePoint * whitePoint)) / (1.0f + L_xy);
            return fToneMappedLum * satpow(f3Color / fInitialPixelLum);
        }
        case DFX_TONE_MAPPING_MODE_UNCHARTED2:
        {
            float  ExposureBias = 2.0f;
            float3 curr         = uncharted2_tonemap(ExposureBias * f3ScaledColor);
            float3 whiteScale   = float3(1.0f, 1.0f, 1.0f) / uncharted2_tonemap(float3(whitePoint, whitePoint, whitePoint));
            return curr * whiteScale;
        }
        case DFX_TONE_MAPPING_MODE_FILMIC_ALU:
        {
            float3 c = max3(f3ScaledColor - float3(0.004f, 0.004f, 0.004f), 0.0f);
            c = (c * (6.2f * c + float3(0.5f, 0.5f, 0.5f))) / (c * (6.2f * c + float3(1.7f, 1.7f, 1.7f)) + float3(0.06f, 0.06f, 0.06f));
            return pow3(c, 2.2f);
        }
        case DFX_TONE_MAPPING_MODE_LOGARITHMIC:
        {
            float fToneMappedLum = std::log10(1.0f + fScaledPixelLum) / std::log10(1.0f + whitePoint);
            return fToneMappedLum * satpow(f3Color / fInitialPixelLum);
        }
        case DFX_TONE_MAPPING_MODE_ADAPTIVE_LOG:
        {
            float Bias = 0.85f;
            float fToneMappedLum = 1.0f / std::log10(1.0f + whitePoint) * std::log(1.0f + fScaledPixelLum) /
                                   std::log(2.0f + 8.0f * std::pow(fScaledPixelLum / whitePoint, std::log(Bias) / std::log(0.5f)));
            return fToneMappedLum * satpow(f3Color / fInitialPixelLum);
        }
        case DFX_TONE_MAPPING_MODE_AGX: return AgXEotf(AgX(f3ScaledColor));
        case DFX_TONE_MAPPING_MODE_AGX_CUSTOM:
        {
            float3 c = AgX(f3ScaledColor);
            c        = AgXPunchyLook(c, At.AgXSaturation, At.AgXOffset, At.AgXSlope, At.AgXPower);
            return AgXEotf(c);
        }
        case DFX_TONE_MAPPING_MODE_PBR_NEUTRAL:
        {
            f3Color = f3Color * (0.3f / fAveLogLum);
            float StartCompression = 0.8f - 0.04f, Desaturation = 0.15f;
            float x      = hmin(f3Color.x, hmin(f3Color.y, f3Color.z));
            float Offset = x < 0.08f ? x - 6.25f * x * x : 0.04f;
            f3Color      = f3Color - float3(Offset, Offset, Offset);
            float Peak   = hmax(f3Color.x, hmax(f3Color.y, f3Color.z));
            if (Peak >= StartCompression)
            {
                float d       = 1.0f - StartCompression;
                float NewPeak = 1.0f - d * d / (Peak + d - StartCompression);
                f3Color       = f3Color * (NewPeak / Peak);
                float g       = 1.0f - 1.0f / (Desaturation * (Peak - NewPeak) + 1.0f);
                f3Color       = lerp(f3Color, float3(NewPeak, NewPeak, NewPeak), g);
            }
            return f3Color;
        }
        case DFX_TONE_MAPPING_MODE_COMMERCE:
        {
            f3Color = f3Color * (0.3f / fAveLogLum);
            float StartCompression = 0.8f, Desaturation = 0.5f;
            float d    = 1.0f - StartCompression;
            float Peak = hmax(f3Color.x, hmax(f3Color.y, f3Color.z));
            if (Peak >= StartCompression)
            {
                float NewPeak = 1.0f - d * d / (Peak + d - StartCompression);
                float InvPeak = 1.0f / Peak;
                float ExtraBrightness = dot(f3Color * (1.0f - StartCompression * InvPeak), float3(1.0f, 1.0f, 1.0f));
                f3Color = f3Color * (NewPeak * InvPeak);
                float g = 1.0f - 3.0f / (Desaturation * ExtraBrightness + 3.0f);
                f3Color = lerp(f3Color, float3(1.0f, 1.0f, 1.0f), g);
            }
            return f3Color;
        }
        default: return f3Color;
    }
}

void tonemap_pass(const dfx_tonemap_attribs& A, float ave_log_lum, bool to_srgb, const TexF4& color, TexF4& out, int threads)
{
    out.resize(color.w, color.h);
    parallel_rows(0, color.h, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < color.w; ++x)
            {
                float4 C = color.load(x, y);
                float3 c = tone_map(C.xyz(), A, ave_log_lum);
                if (to_srgb) c = LinearToSRGB(c);
                out.at(x, y) = float4(c, C.w);
            }
    });
}

} // namespace orc
