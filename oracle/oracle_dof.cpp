// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h). Pinned against the reference's own shaders (oracle/refshader, tests/test_reference_shaders.py).
// DepthOfField, eleven passes restated from Shaders/PostProcess/DepthOfField/private/DOF_*.fx (file:line per
// function) in the order of DepthOfField::Execute (PostProcess/DepthOfField/src/DepthOfField.cpp:292-331). Storage is fp32 like
// the rest of the oracle (the reference keeps CoC in R16_FLOAT / R16_UNORM and colour in RGBA16_FLOAT / R11G11B10_FLOAT).
#include "oracle.h"

namespace orc
{

// DepthOfField.cpp:49-73 (https://www.shadertoy.com/view/wdKXDK): rings from the outside in, ring i has max(density * i, 1) points
std::vector<float2> dof_kernel_points(int RingCount, int RingDensity)
{
    std::vector<float2> Kernel;
    float RadiusInc = 1.0f / (static_cast<float>(RingCount) - 1.0f);
    for (int i = RingCount - 1; i >= 0; --i)
    {
        int   PointCount = std::max(RingDensity * i, 1);
        float Radius     = static_cast<float>(i) * RadiusInc;
        float ThetaInc   = 2.0f * 3.14159265358979323846f / static_cast<float>(PointCount);
        float Offset     = 0.1f * static_cast<float>(i);
        for (int j = 0; j < PointCount; ++j)
        {
            float Theta = Offset + static_cast<float>(j) * ThetaInc;
            Kernel.push_back(float2(std::cos(Theta), std::sin(Theta)) * Radius);
        }
    }
    return Kernel;
}

// DepthOfField.cpp:75-91
std::vector<float> dof_gauss_kernel(int Radius, float Sigma)
{
    std::vector<float> Kernel;
    float              Sum = 0.0f;
    for (int i = -Radius; i <= Radius; ++i)
    {
        float Value = std::exp(-static_cast<float>(i * i) / (2.0f * Sigma * Sigma));
        Kernel.push_back(Value);
        Sum += Value;
    }
    for (float& Value : Kernel) Value /= Sum;
    return Kernel;
}

static inline int ComputeSampleCount(int RingCount, int RingDensity) { return 1 + RingDensity * ((RingCount - 1) * RingCount >> 1); } // DOF_Common.fx:4-7
static inline float ComputeHDRWeight(float3 Color) { return 1.0f + Luminance(Color); }                                                // :9-12
static inline float ComputeSDRWeight(float3 Color) { return 1.0f / (1.0f + Luminance(Color)); }                                       // :14-17

// D1  DOF_ComputeCircleOfConfusion.fx:24-39
void dof_circle_of_confusion(const Camera& cam, const dfx_dof_attribs& A, const TexF& depth, TexF& coc, int threads)
{
    coc.resize(depth.w, depth.h);
    parallel_rows(0, depth.h, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < depth.w; ++x)
            {
                float LinearDepth = DepthToCameraZ(depth.load(x, y), cam.mProj);
                float f   = cam.fFocalLength / 1000.0f;
                float K   = f * f / (cam.fFStop * (cam.fFocusDistance - f));
                float CoC = K * (LinearDepth - cam.fFocusDistance) / hmax(LinearDepth, 1e-4f);
                coc.at(x, y) = clampf(1000.0f * CoC / (cam.fSensorWidth * A.MaxCircleOfConfusion), -1.0f, 1.0f);
            }
    });
}

// D2  DOF_ComputeTemporalCircleOfConfusion.fx:54-92 (g_TextureCurrCoC: point clamp, g_TexturePrevCoC: linear clamp, …cpp:445-446)
void dof_temporal_coc(const Camera& cam, const dfx_dof_attribs& A, const TexF& curr, const TexF& prev, const TexF2& closest_motion, TexF& out, int threads)
{
    const int W = curr.w, H = curr.h;
    out.resize(W, H);
    const float2 Viewport(cam.f4ViewportSize.x, cam.f4ViewportSize.y), InvViewport(cam.f4ViewportSize.z, cam.f4ViewportSize.w);
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                float2 Motion       = closest_motion.load(px, py) * float2(0.5f, -0.5f); // F3NDC_XYZ_TO_UVD_SCALE.xy
                float2 PrevPosition = Position - Motion * Viewport;
                if (!IsInsideScreen(PrevPosition, Viewport))
                {
                    out.at(px, py) = curr.load(px, py);
                    continue;
                }
                float CoCCurr = curr.load(px, py);
                float CoCPrev = sample_linear(prev, PrevPosition * InvViewport, Address::Clamp);
                float M1 = 0.0f, M2 = 0.0f;
                for (int x = -1; x <= 1; x++)
                    for (int y = -1; y <= 1; y++)
                    {
                        float2 Location = Position + float2(float(x), float(y));
                        float  CoC      = sample_point_clamp(curr, Location * InvViewport);
                        M1 += CoC;
                        M2 += CoC * CoC;
                    }
                float Mean = M1 / 9.0f, Variance = (M2 / 9.0f) - (Mean * Mean), StdDev = std::sqrt(hmax(Variance, 0.0f));
                float CoCMin = Mean - 2.5f * StdDev, CoCMax = Mean + 2.5f * StdDev; // DOF_TEMPORAL_VARIANCE_GAMMA
                out.at(px, py) = lerp(CoCCurr, clampf(CoCPrev, CoCMin, CoCMax), A.TemporalStabilityFactor);
            }
    });
}

// D3  DOF_ComputeSeparatedCircleOfConfusion.fx:5-11 : the near field, |CoC| where CoC < 0
void dof_separated_coc(const TexF& coc, TexF& out, int threads)
{
    out.resize(coc.w, coc.h);
    parallel_rows(0, coc.h, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < coc.w; ++x)
            {
                float CoC    = coc.load(x, y);
                out.at(x, y) = std::fabs(CoC) * float(CoC < 0.0f);
            }
    });
}

// D4  DOF_ComputeDilationCircleOfConfusion.fx:16-52 : 2x2 (+ odd row / column) maximum, dimensions (w >> 1, h >> 1)
void dof_dilation_level(const TexF& last, TexF& out, int threads)
{
    out.resize(last.w >> 1, last.h >> 1);
    const bool IsWidthOdd = (last.w & 1) != 0, IsHeightOdd = (last.h & 1) != 0;
    parallel_rows(0, out.h, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < out.w; ++x)
            {
                auto  S      = [&](int ox, int oy) { return last.load_clamped(2 * x + ox, 2 * y + oy); };
                float MaxCoC = hmax(hmax(S(0, 0), S(0, 1)), hmax(S(1, 0), S(1, 1)));
                if (IsWidthOdd) MaxCoC = hmax(MaxCoC, hmax(S(2, 0), S(2, 1)));
                if (IsHeightOdd) MaxCoC = hmax(MaxCoC, hmax(S(0, 2), S(1, 2)));
                if (IsWidthOdd && IsHeightOdd) MaxCoC = hmax(MaxCoC, S(2, 2));
                out.at(x, y) = MaxCoC;
            }
    });
}

// D5, D6  DOF_ComputeBlurredCircleOfConfusion.fx:8-28 (13-tap Gaussian, radius 6, sigma 5, clamped addressing)
void dof_blur_coc(const TexF& coc, bool vertical, TexF& out, int threads)
{
    const std::vector<float> Gauss = dof_gauss_kernel(6, 5.0f);
    out.resize(coc.w, coc.h);
    parallel_rows(0, coc.h, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < coc.w; ++x)
            {
                float ResultSum = 0.0f;
                for (int i = -6; i <= 6; i++)
                {
                    float CoC = vertical ? coc.load_clamped(x, y + i) : coc.load_clamped(x + i, y);
                    ResultSum += CoC * Gauss[size_t(i + 6)];
                }
                out.at(x, y) = ResultSum;
            }
    });
}

// D7  DOF_ComputePrefilteredTexture.fx:23-52 : half-resolution colour (SDR-weighted mean of the 2x2 block) with the near alpha from
// the blurred dilation texture (linear clamp) and the far alpha = max CoC of the block if positive
void dof_prefilter(const TexF4& color, const TexF& coc, const TexF& dilation, TexF4& out_fg, TexF4& out_bg, int threads)
{
    const int W = color.w / 2, H = color.h / 2;
    out_fg.resize(W, H), out_bg.resize(W, H);
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Texcoord((float(px) + 0.5f) / float(W), (float(py) + 0.5f) / float(H));
                float  CoCMax = -FLT_MAX_F;
                float4 ColorSum;
                for (int i = 0; i < 4; ++i)
                {
                    int    lx = 2 * px + (i & 1), ly = 2 * py + (i >> 1);
                    float3 Color  = color.load(lx, ly).xyz();
                    float  CoC    = coc.load(lx, ly);
                    float  Weight = ComputeSDRWeight(Color);
                    CoCMax        = hmax(CoCMax, CoC);
                    ColorSum      = ColorSum + float4(Color, 1.0f) * Weight;
                }
                float  ForegroundAlpha = sample_linear(dilation, Texcoord, Address::Clamp);
                float  BackgroundAlpha = std::fabs(CoCMax) * float(CoCMax > 0.0f);
                float3 Mean            = ColorSum.xyz() / hmax(ColorSum.w, 1.e-5f);
                out_fg.at(px, py)      = float4(Mean, ForegroundAlpha);
                out_bg.at(px, py)      = float4(Mean, BackgroundAlpha);
            }
    });
}

// D8  DOF_ComputeBokehFirstPass.fx:47-104 : gather over the Octaweb kernel, radius = 0.5 * CoC * MaxCircleOfConfusion in UV
void dof_bokeh_first(const Camera& cam, const dfx_dof_attribs& A, uint flags, const TexF4& fg, const TexF4& bg, const TexF4& radiance, TexF4& out_fg,
                     TexF4& out_bg, int threads)
{
    const int W = fg.w, H = fg.h;
    out_fg.resize(W, H), out_bg.resize(W, H);
    const std::vector<float2> Kernel = dof_kernel_points(A.BokehKernelRingCount, A.BokehKernelRingDensity);
    const int   SampleCount = ComputeSampleCount(A.BokehKernelRingCount, A.BokehKernelRingDensity);
    const float AspectRatio = cam.f4ViewportSize.x * cam.f4ViewportSize.w;
    const bool  Karis       = (flags & 2u) != 0; // FEATURE_FLAG_ENABLE_KARIS_INVERSE
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Center((float(px) + 0.5f) / float(W), (float(py) + 0.5f) / float(H));
                float  CoCNear = sample_linear(fg, Center, Address::Clamp).w, CoCFar = sample_linear(bg, Center, Address::Clamp).w;
                float4 Fore, Back;
                if (CoCNear > 0.0f)
                    for (int i = 0; i < SampleCount; i++)
                    {
                        float2 SamplePosition = Kernel[size_t(i)] * 0.5f * CoCNear * A.MaxCircleOfConfusion;
                        float2 Offset(SamplePosition.x, AspectRatio * SamplePosition.y);
                        float4 S      = sample_linear(fg, Center + Offset, Address::Clamp);
                        float  Weight = Karis ? ComputeHDRWeight(sample_linear(radiance, Center + Offset, Address::Clamp).xyz()) : 1.0f;
                        Fore          = Fore + float4(S.xyz(), 1.0f) * Weight;
                    }
                if (CoCFar > 0.0f)
                    for (int i = 0; i < SampleCount; i++)
                    {
                        float2 SamplePosition = Kernel[size_t(i)] * 0.5f * CoCFar * A.MaxCircleOfConfusion;
                        float2 Offset(SamplePosition.x, AspectRatio * SamplePosition.y);
                        float4 S      = sample_linear(bg, Center + Offset, Address::Clamp);
                        float  Weight = Karis ? ComputeHDRWeight(sample_linear(radiance, Center + Offset, Address::Clamp).xyz()) : 1.0f;
                        Back          = Back + float4(S.xyz(), 1.0f) * Weight * float(S.w >= CoCFar);
                    }
                out_fg.at(px, py) = float4(Fore.xyz() * rcp(Fore.w + float(Fore.w == 0.0f)), CoCNear);
                out_bg.at(px, py) = float4(Back.xyz() * rcp(Back.w + float(Back.w == 0.0f)), CoCFar);
            }
    });
}

// D9  DOF_ComputeBokehSecondPass.fx:37-85 : flood fill with the small kernel (3 rings x 5), component-wise maximum
void dof_bokeh_second(const Camera& cam, const dfx_dof_attribs& A, const TexF4& fg, const TexF4& bg, TexF4& out_fg, TexF4& out_bg, int threads)
{
    const int W = fg.w, H = fg.h;
    out_fg.resize(W, H), out_bg.resize(W, H);
    const std::vector<float2> Kernel = dof_kernel_points(3, 5);
    const int   SampleCount = ComputeSampleCount(3, 5);
    const float AspectRatio = cam.f4ViewportSize.x * cam.f4ViewportSize.w;
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Center((float(px) + 0.5f) / float(W), (float(py) + 0.5f) / float(H));
                float4 Fore = sample_linear(fg, Center, Address::Clamp), Back = sample_linear(bg, Center, Address::Clamp);
                float  CoCNear = Fore.w, CoCFar = Back.w;
                float3 F = Fore.xyz(), B = Back.xyz();
                if (CoCNear > 0.0f)
                    for (int i = 0; i < SampleCount; i++)
                    {
                        float2 SamplePosition = Kernel[size_t(i)] * 0.25f * CoCNear * A.MaxCircleOfConfusion;
                        float4 S = sample_linear(fg, Center + float2(SamplePosition.x, AspectRatio * SamplePosition.y), Address::Clamp);
                        F        = float3(hmax(S.x, F.x), hmax(S.y, F.y), hmax(S.z, F.z));
                    }
                if (CoCFar > 0.0f)
                    for (int i = 0; i < SampleCount; i++)
                    {
                        float2 SamplePosition = Kernel[size_t(i)] * 0.25f * CoCFar * A.MaxCircleOfConfusion;
                        float4 S = sample_linear(bg, Center + float2(SamplePosition.x, AspectRatio * SamplePosition.y), Address::Clamp);
                        float  k = float(S.w >= CoCFar);
                        B        = float3(hmax(S.x * k, B.x), hmax(S.y * k, B.y), hmax(S.z * k, B.z));
                    }
                out_fg.at(px, py) = float4(F, CoCNear);
                out_bg.at(px, py) = float4(B, CoCFar);
            }
    });
}

// D10  DOF_ComputePostfilteredTexture.fx:26-48 : 2x2 tent (four bilinear taps at +-half a texel)
void dof_postfilter(const TexF4& fg, const TexF4& bg, TexF4& out_fg, TexF4& out_bg, int threads)
{
    const int W = fg.w, H = fg.h;
    out_fg.resize(W, H), out_bg.resize(W, H);
    const float2 TexelSize(rcp(float(W)), rcp(float(H)));
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Center((float(px) + 0.5f) / float(W), (float(py) + 0.5f) / float(H));
                auto   tap = [&](const TexF4& t, float ox, float oy) { return sample_linear(t, Center + TexelSize * float2(ox, oy), Address::Clamp); };
                out_fg.at(px, py) = (tap(fg, -0.5f, -0.5f) + tap(fg, -0.5f, +0.5f) + tap(fg, +0.5f, -0.5f) + tap(fg, +0.5f, +0.5f)) * 0.25f;
                out_bg.at(px, py) = (tap(bg, -0.5f, -0.5f) + tap(bg, -0.5f, +0.5f) + tap(bg, +0.5f, -0.5f) + tap(bg, +0.5f, +0.5f)) * 0.25f;
            }
    });
}

// D11  DOF_ComputeCombinedTexture.fx:34-46 : far field, then near field over the full-resolution source
void dof_combine(const dfx_dof_attribs& A, const TexF4& color, const TexF4& dof_near, const TexF4& dof_far, TexF4& out, int threads)
{
    const int W = color.w, H = color.h;
    out.resize(W, H);
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Texcoord((float(px) + 0.5f) / float(W), (float(py) + 0.5f) / float(H));
                float4 Src = color.load(px, py);
                float3 SourceFullRes = Src.xyz();
                float4 DoFNear = sample_linear(dof_near, Texcoord, Address::Clamp), DoFFar = sample_linear(dof_far, Texcoord, Address::Clamp);
                float3 Result = lerp(SourceFullRes, DoFFar.xyz(), smoothstep(0.1f, 1.0f, DoFFar.w));
                Result        = lerp(Result, DoFNear.xyz(), smoothstep(0.1f, 1.0f, DoFNear.w));
                out.at(px, py) = float4(lerp(SourceFullRes, Result, A.AlphaInterpolation), Src.w); // the target is RGB; alpha kept for the next pass
            }
    });
}

} // namespace orc
