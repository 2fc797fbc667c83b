// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h). Pinned against the reference's own shaders (oracle/refshader, tests/test_reference_shaders.py).
// PostFXContext passes P0-P2 and SSAO passes A1-A8, restated from the reference HLSL (file:line cited per function).
#include "oracle.h"

namespace orc
{

// =====================================================================================================================
// PostFXContext
// =====================================================================================================================

// ComputeBlueNoiseTexture.fx:20-32 SampleRandomNumber
static float SampleRandomNumber(const uint8_t* tables, uint px, uint py, uint SampleDimension)
{
    const uint8_t* Sobol = tables;
    const uint8_t* Tile  = tables + 256;
    px &= 127u;
    py &= 127u;
    SampleDimension &= 255u;
    uint Value         = Sobol[SampleDimension];
    uint OriginalIndex = (SampleDimension % 8u) + (px + py * 128u) * 8u;
    // g_ScramblingTileBuffer is a 512x256 R8_UINT texture addressed (idx % 512, idx / 512) == linear index
    Value = Value ^ uint(Tile[OriginalIndex]);
    return (float(Value) + 0.5f) / 256.0f;
}

// ComputeBlueNoiseTexture.fx:34-57 HilbertIndex (HILBERT_LEVEL 7)
static uint HilbertIndex(uint px, uint py)
{
    const uint W = 128u;
    px &= (W - 1u);
    py &= (W - 1u);
    uint Index = 0u;
    for (uint CurLevel = W / 2u; CurLevel > 0u; CurLevel /= 2u)
    {
        uint RegionX = uint((px & CurLevel) > 0u);
        uint RegionY = uint((py & CurLevel) > 0u);
        Index += CurLevel * CurLevel * ((3u * RegionX) ^ RegionY);
        if (RegionY == 0u)
        {
            if (RegionX == 1u)
            {
                px = (W - 1u) - px;
                py = (W - 1u) - py;
            }
            uint Temp = px;
            px        = py;
            py        = Temp;
        }
    }
    return Index;
}

void postfx_blue_noise(const uint8_t* tables, uint FrameIndex, TexF2& xy, TexF2& zw)
{
    xy.resize(128, 128);
    zw.resize(128, 128);
    for (uint y = 0; y < 128; ++y)
        for (uint x = 0; x < 128; ++x)
        {
            // SampleRandomVector2D :60-68 (R1 sequence)
            float G     = 1.61803398875f;
            float Alpha = 0.5f + rcp(G) * float(FrameIndex & 0xFFu);
            xy.at(x, y) = float2(frac(SampleRandomNumber(tables, x, y, 0u) + Alpha), frac(SampleRandomNumber(tables, x, y, 1u) + Alpha));
            // SampleRandomVector1D1D :71-79 (R2 sequence over the Hilbert index)
            uint Index = HilbertIndex(x, y) + FrameIndex;
            Index += 288u * (FrameIndex & 127u);
            float  G2 = 1.32471795724474602596f;
            float2 A2 = float2(rcp(G2), rcp(G2 * G2));
            zw.at(x, y) = float2(frac(0.5f + float(Index) * A2.x), frac(0.5f + float(Index) * A2.y));
        }
}

void postfx_reprojected_depth(const Camera& curr, const Camera& prev, const TexF& depth, TexF& out, int threads)
{
    out.resize(depth.w, depth.h);
    parallel_rows(0, depth.h, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < depth.w; ++x)
            {
                float2 Position(float(x) + 0.5f, float(y) + 0.5f);
                float  Depth = depth.load(x, y);
                float3 CurrScreenCoord(Position * float2(curr.f4ViewportSize.z, curr.f4ViewportSize.w), Depth);
                CurrScreenCoord.x += F3NDC_XYZ_TO_UVD_SCALE.x * curr.f2Jitter.x;
                CurrScreenCoord.y += F3NDC_XYZ_TO_UVD_SCALE.y * curr.f2Jitter.y;
                float3 WorldPosition   = InvProjectPosition(CurrScreenCoord, curr.mViewProjInv);
                float3 PrevScreenCoord = ProjectPosition(WorldPosition, prev.mViewProj);
                out.at(x, y)           = PrevScreenCoord.z;
            }
    });
}

void postfx_closest_motion(const TexF& depth, const TexF2& motion, TexF2& out, int threads)
{
    out.resize(depth.w, depth.h);
    parallel_rows(0, depth.h, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < depth.w; ++px)
            {
                float ClosestDepth = g_reversed_depth ? 0.0f : 1.0f; // DepthFarPlane (ComputeClosestMotion.fx:5-9)
                int2  ClosestOffset(0, 0);
                for (int x = -1; x <= 1; x++)
                    for (int y = -1; y <= 1; y++)
                    {
                        float NeighborDepth = depth.load(px + x, py + y); // unclamped Load: OOB -> 0
                        if (g_reversed_depth ? NeighborDepth > ClosestDepth : NeighborDepth < ClosestDepth) // :36-40
                        {
                            ClosestOffset = int2(x, y);
                            ClosestDepth  = NeighborDepth;
                        }
                    }
                out.at(px, py) = motion.load(px + ClosestOffset.x, py + ClosestOffset.y);
            }
    });
}

// =====================================================================================================================
// SSAO
// =====================================================================================================================

static inline bool IsBackground(float Depth) { return g_reversed_depth ? Depth < 1e-6f : Depth >= (1.0f - 1e-6f); } // SSAO_Common.fxh:16-23

// SSAO_Common.fxh:25-28
static inline float ComputeGeometryWeight(float3 CenterPos, float3 TapPos, float3 CenterNormal, float PlaneDistanceNorm)
{
    return saturate(1.0f - std::fabs(dot((TapPos - CenterPos), CenterNormal)) * PlaneDistanceNorm);
}

// SSAO_ComputePrefilteredDepthBuffer.fx:42-71
static float ComputeDepthMIPFiltered(const dfx_ssao_attribs& A, const float* SampledDepth, uint Count)
{
    float WeightDepth = SampledDepth[0];
    for (uint Idx = 1u; Idx < Count; Idx++) WeightDepth = hmin(WeightDepth, SampledDepth[Idx]);

    float DepthRangeScaleFactor = 0.75f;
    float EffectRadius          = DepthRangeScaleFactor * A.EffectRadius * A.RadiusMultiplier;
    float FalloffRange          = A.EffectFalloffRange * EffectRadius;
    float FalloffFrom           = EffectRadius - FalloffRange;
    float FalloffMul            = -1.0f / (FalloffRange);
    float FalloffAdd            = FalloffFrom / FalloffRange + 1.0f;

    float DepthSum = 0.0f, WeightSum = 0.0f;
    for (uint Idx = 0u; Idx < Count; Idx++)
    {
        float Weight = saturate(std::fabs(WeightDepth - SampledDepth[Idx]) * FalloffMul + FalloffAdd);
        DepthSum += Weight * SampledDepth[Idx];
        WeightSum += Weight;
    }
    return DepthSum / WeightSum;
}

// Gathers the 2x2 (+ odd row/column) footprint in the order of the reference's ArrayAppend calls.
template <class F>
static uint gather_footprint(const TexF& last, int rx, int ry, F&& xform, float* out)
{
    auto ld = [&](int ox, int oy) { return xform(last.load_clamped(rx + ox, ry + oy)); };
    uint n  = 0;
    out[n++] = ld(0, 0);
    out[n++] = ld(0, 1);
    out[n++] = ld(1, 0);
    out[n++] = ld(1, 1);
    bool IsWidthOdd = (last.w & 1) != 0, IsHeightOdd = (last.h & 1) != 0;
    if (IsWidthOdd)
    {
        out[n++] = ld(2, 0);
        out[n++] = ld(2, 1);
    }
    if (IsHeightOdd)
    {
        out[n++] = ld(0, 2);
        out[n++] = ld(1, 2);
    }
    if (IsWidthOdd && IsHeightOdd) out[n++] = ld(2, 2);
    return n;
}

void ssao_prefilter_depth(const Camera& cam, const dfx_ssao_attribs& A, const TexF& depth, MipTex<float>& pyr, int threads)
{
    const int levels = std::min(compute_mip_levels_count(depth.w, depth.h), 5); // SSAO_DEPTH_PREFILTERED_MAX_MIP + 1
    pyr.create(depth.w, depth.h, levels);
    pyr.mip[0].d = depth.d; // CopyTextureDepth
    for (int m = 1; m < levels; ++m)
    {
        const TexF& last = pyr.mip[m - 1];
        TexF&       dst  = pyr.mip[m];
        parallel_rows(0, dst.h, threads, [&](int ya, int yb) {
            for (int y = ya; y < yb; ++y)
                for (int x = 0; x < dst.w; ++x)
                {
                    float s[9];
                    uint  n = gather_footprint(last, 2 * x, 2 * y, [&](float d) { return DepthToCameraZ(d, cam.mProj); }, s);
                    dst.at(x, y) = saturate(CameraZToDepth(ComputeDepthMIPFiltered(A, s, n), cam.mProj));
                }
        });
    }
}

// SSAO_ComputeAmbientOcclusion.fx:47-53
static inline float FastACos(float Value)
{
    float AbsValue = std::fabs(Value);
    float Result   = -0.156583f * AbsValue + M_HALF_PI_F;
    Result *= std::sqrt(1.0f - AbsValue);
    return (Value >= 0.0f) ? Result : M_PI_F - Result;
}
float oracle_fast_acos(float v) { return FastACos(v); }

// :55-58
static inline float IntegrateArcUniform(float HorizonX, float HorizonY) { return (1.0f - std::cos(HorizonX) + (1.0f - std::cos(HorizonY))); }
// :60-66
static inline float IntegrateArcCosWeighted(float HorizonX, float HorizonY, float N, float CosN)
{
    float H1 = HorizonX * 2.0f, H2 = HorizonY * 2.0f;
    float SinN = std::sin(N);
    return 0.25f * ((-std::cos(H1 - N) + CosN + H1 * SinN) + (-std::cos(H2 - N) + CosN + H2 * SinN));
}

// :77-99
static uint ComputeOccludedSectors(float MinHorizon, float MaxHorizon, uint OccludedBitfield)
{
    MinHorizon  = saturate(MinHorizon);
    MaxHorizon  = saturate(MaxHorizon);
    uint Result = OccludedBitfield;
    if (MaxHorizon > MinHorizon)
    {
        uint SectorCount = 32u;
        uint StartInt    = std::min(uint(MinHorizon * float(SectorCount)), SectorCount - 1u);
        uint EndInt      = std::min(uint(std::ceil(MaxHorizon * float(SectorCount))), SectorCount);
        if (EndInt > StartInt)
        {
            uint AngleInt      = EndInt - StartInt;
            uint AngleBitfield = AngleInt >= 32u ? 0xFFFFFFFFu : ((1u << AngleInt) - 1u);
            Result |= AngleBitfield << StartInt;
        }
    }
    return Result;
}

// :101-119
static uint ComputeSampleOcclusion(const dfx_ssao_attribs& A, float3 S0, float3 S1, float3 PositionVS, float3 ViewVS, float NSlice,
                                   float FalloffMul, float FalloffAdd, uint OccludedBitfield)
{
    float3 DeltaPos0     = S0 - PositionVS;
    float3 DeltaPos1     = S1 - PositionVS;
    float3 ViewThickness = ViewVS * A.BitmaskThickness;
    float2 Weight        = saturate(float2(length(DeltaPos0), length(DeltaPos1)) * FalloffMul + float2(FalloffAdd, FalloffAdd));
    float4 FrontBack(FastACos(dot(normalize(DeltaPos0), ViewVS)), FastACos(dot(normalize(DeltaPos0 - ViewThickness), ViewVS)),
                     FastACos(dot(normalize(DeltaPos1), ViewVS)), FastACos(dot(normalize(DeltaPos1 - ViewThickness), ViewVS)));
    FrontBack = saturate((float4(-FrontBack.x, -FrontBack.y, FrontBack.z, FrontBack.w) - float4(NSlice, NSlice, NSlice, NSlice) +
                          float4(M_HALF_PI_F, M_HALF_PI_F, M_HALF_PI_F, M_HALF_PI_F)) /
                         M_PI_F);
    if (Weight.x > 0.0f) OccludedBitfield = ComputeOccludedSectors(FrontBack.y, FrontBack.x, OccludedBitfield);
    if (Weight.y > 0.0f) OccludedBitfield = ComputeOccludedSectors(FrontBack.z, FrontBack.w, OccludedBitfield);
    return OccludedBitfield;
}

// :121-130
static float2 ComputeSampleHorizons(float3 S0, float3 S1, float3 PositionVS, float3 ViewVS, float2 MinCosHorizons, float2 MaxCosHorizons,
                                    float FalloffMul, float FalloffAdd)
{
    float3 D0 = S0 - PositionVS, D1 = S1 - PositionVS;
    float2 SampleDistance(length(D0), length(D1));
    float2 SampleCosHorizon(dot(D0 / SampleDistance.x, ViewVS), dot(D1 / SampleDistance.y, ViewVS));
    float2 Weight = saturate(SampleDistance * FalloffMul + float2(FalloffAdd, FalloffAdd));
    float2 l(lerp(MinCosHorizons.x, SampleCosHorizon.x, Weight.x), lerp(MinCosHorizons.y, SampleCosHorizon.y, Weight.y));
    return float2(hmax(MaxCosHorizons.x, l.x), hmax(MaxCosHorizons.y, l.y));
}

// A0 SSAO_ComputeDownsampledDepth.fx:8-29
void ssao_downsample_depth(const TexF& depth, TexF& out, int threads)
{
    const int W = depth.w / 2, H = depth.h / 2;
    out.resize(W, H);
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < W; ++x)
            {
                float Depth0 = depth.load(2 * x + 0, 2 * y + 0), Depth1 = depth.load(2 * x + 0, 2 * y + 1);
                float Depth2 = depth.load(2 * x + 1, 2 * y + 0), Depth3 = depth.load(2 * x + 1, 2 * y + 1);
                float MinDepth = hmin(hmin(Depth0, Depth1), hmin(Depth2, Depth3));
                float MaxDepth = hmax(hmax(Depth0, Depth1), hmax(Depth2, Depth3));
                int   Pattern  = ((x + y) & 1) & 1; // ComputeCheckerboardPattern :8-11
                out.at(x, y)   = lerp(MinDepth, MaxDepth, float(Pattern));
            }
    });
}

// A4 SSAO_ComputeBilateralUpsampling.fx:62-139 (g_TextureDepth: linear clamp, g_TextureOcclusion: linear clamp, …cpp:614-615)
void ssao_bilateral_upsampling(const Camera& cam, const TexF& depth, const TexF& occ, TexF& out, int threads)
{
    const int W = depth.w, H = depth.h;
    out.resize(W, H);
    const float2 InvViewport(cam.f4ViewportSize.z, cam.f4ViewportSize.w);
    const int2   HalfDim(int(0.5f * cam.f4ViewportSize.x), int(0.5f * cam.f4ViewportSize.y));
    const float  Sigma = 0.9f, DepthSigma = 0.0075f; // SSAO_BILATERAL_UPSAMPLING_SIGMA / _DEPTH_SIGMA (Structures.fxh:38,41)
    auto ComputeDepthWeight = [&](float CenterDepth, float GuideDepth, float S) {
        float LinearDepth0 = DepthToCameraZ(CenterDepth, cam.mProj);
        float LinearDepth1 = DepthToCameraZ(GuideDepth, cam.mProj);
        float Alpha        = std::fabs(LinearDepth0 - LinearDepth1) / hmax(LinearDepth0, 1e-6f);
        return std::exp(-(Alpha * Alpha) / (2.0f * S * S));
    };
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float CenterDepth = depth.load(px, py);
                if (IsBackground(CenterDepth))
                {
                    out.at(px, py) = 1.0f;
                    continue;
                }
                int2  CenterLocation(int(0.5f * float(px)), int(0.5f * float(py))); // int2(0.5 * floor(Position.xy))
                float OcclusionSum = 0.0f, WeightSum = 0.0f;
                for (int x = -1; x <= 1; x++)
                    for (int y = -1; y <= 1; y++)
                    {
                        int2   Location = ClampScreenCoord(int2(CenterLocation.x + x, CenterLocation.y + y), HalfDim);
                        float2 Texcoord = float2(2.0f * (float(Location.x) + 0.5f), 2.0f * (float(Location.y) + 0.5f)) * InvViewport;
                        float  SampledSignal = occ.load(Location.x, Location.y);
                        float  SampledGuided = sample_linear(depth, Texcoord, Address::Clamp);
                        float  WeightS       = ComputeSpatialWeight(float(x * x + y * y), Sigma);
                        float  WeightZ       = ComputeDepthWeight(CenterDepth, SampledGuided, DepthSigma);
                        OcclusionSum += WeightS * WeightZ * SampledSignal;
                        WeightSum += WeightS * WeightZ;
                    }
                float2 CenterUV = float2(2.0f * (float(CenterLocation.x) + 0.5f), 2.0f * (float(CenterLocation.y) + 0.5f)) * InvViewport;
                out.at(px, py)  = WeightSum > 0.0f ? OcclusionSum / WeightSum : sample_linear(occ, CenterUV, Address::Clamp);
            }
    });
}

void ssao_ambient_occlusion(const Camera& cam, const dfx_ssao_attribs& A, const MipTex<float>& pre, const TexF4& normal,
                            const TexF2& blue_noise_zw, TexF& out, int threads, bool half_res, bool half_precision_depth)
{
    const int W = pre.mip[0].w, H = pre.mip[0].h;
    out.resize(W, H, 1.0f); // ClearRenderTarget 1.0
    const float  ivs = half_res ? 2.0f : 1.0f; // GetInvViewportSize() :68-75
    const float2 InvViewport(ivs * cam.f4ViewportSize.z, ivs * cam.f4ViewportSize.w);
    const float2 Viewport(cam.f4ViewportSize.x, cam.f4ViewportSize.y);

    auto SamplePrefilteredDepth = [&](float2 uv, float MipLevel) {
        return sample_point_clamp(pre.mip[nearest_mip(MipLevel, pre.levels())], uv);
    };

    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                float2 ScreenCoordUV = Position * InvViewport;
                float3 PositionSS(ScreenCoordUV, SamplePrefilteredDepth(ScreenCoordUV, 0.0f));
                if (IsBackground(PositionSS.z)) continue; // discard

                float3 NormalVS   = mul_dir(sample_point_clamp(normal, ScreenCoordUV).xyz(), cam.mView);
                float3 PositionVS = ScreenXYDepthToViewSpace(PositionSS, cam.mProj);
                float  Offset     = half_precision_depth ? 0.005f : 0.00001f; // :145-150
                PositionVS        = PositionVS + NormalVS * Offset * PositionVS.z;

                float3 ViewVS = -normalize(PositionVS);
                float2 Xi     = blue_noise_zw.load(px & 127, py & 127);

                float EffectRadius = A.EffectRadius * A.RadiusMultiplier;
                float FalloffRange = A.EffectFalloffRange * EffectRadius;
                float FalloffFrom  = EffectRadius - FalloffRange;
                float FalloffMul   = -1.0f / FalloffRange;
                float FalloffAdd   = FalloffFrom / FalloffRange + 1.0f;
                float SampleRadius = 0.5f * EffectRadius * cam.mProj.m[0][0];
                if (cam.mProj.m[3][3] == 0.0f) SampleRadius /= PositionVS.z;

                float Visibility = 0.0f;
                for (int SliceIdx = 0; SliceIdx < 3; SliceIdx++)
                {
                    // ComputeSliceDirection :40-45
                    float  Rotation = float(SliceIdx) / 3.0f;
                    float  Phi      = (Xi.x + Rotation) * M_PI_F;
                    float2 Omega(std::cos(Phi), std::sin(Phi));

                    float3 SliceDirection(Omega, 0.0f);
                    float3 OrthoSliceDir = SliceDirection - dot(SliceDirection, ViewVS) * ViewVS;
                    float3 Axis          = normalize(cross(SliceDirection, ViewVS));
                    float3 ProjNormal    = NormalVS - Axis * dot(NormalVS, Axis);

                    float ProjNormalLen = length(ProjNormal);
                    float CosNorm       = saturate(dot(ProjNormal / ProjNormalLen, ViewVS));
                    float N             = sign(dot(OrthoSliceDir, ProjNormal)) * FastACos(CosNorm);

                    uint   OccludedBitfield = 0u;
                    float  NBitmask         = -N;
                    float2 MinCosHorizons(std::cos(N + M_HALF_PI_F), std::cos(N - M_HALF_PI_F));
                    float2 MaxCosHorizons = MinCosHorizons;

                    float2 SampleDirection = float2(Omega.x, Omega.y) * float2(F3NDC_XYZ_TO_UVD_SCALE.x, F3NDC_XYZ_TO_UVD_SCALE.y) * SampleRadius;
                    SampleDirection.x *= cam.f4ViewportSize.y * cam.f4ViewportSize.z;

                    for (int SampleIdx = 0; SampleIdx < 3; SampleIdx++)
                    {
                        float  Noise  = frac(Xi.y + float(SliceIdx + SampleIdx * 3) * 0.6180339887498948482f);
                        float  Sample = (float(SampleIdx) + Noise) / 3.0f;
                        float2 SampleOffset = Sample * Sample * SampleDirection;
                        float2 SS0 = PositionSS.xy() + SampleOffset;
                        float2 SS1 = PositionSS.xy() - SampleOffset;

                        float MipLevel = clampf(std::log2(length(SampleOffset * Viewport)) - A.DepthMIPSamplingOffset, 0.0f, 4.0f);
                        float3 VS0 = ScreenXYDepthToViewSpace(float3(SS0, SamplePrefilteredDepth(SS0, MipLevel)), cam.mProj);
                        float3 VS1 = ScreenXYDepthToViewSpace(float3(SS1, SamplePrefilteredDepth(SS1, MipLevel)), cam.mProj);

                        if (A.Algorithm == DFX_SSAO_ALGORITHM_VBAO)
                            OccludedBitfield = ComputeSampleOcclusion(A, VS0, VS1, PositionVS, ViewVS, NBitmask, FalloffMul, FalloffAdd, OccludedBitfield);
                        else
                            MaxCosHorizons = ComputeSampleHorizons(VS0, VS1, PositionVS, ViewVS, MinCosHorizons, MaxCosHorizons, FalloffMul, FalloffAdd);
                    }

                    if (A.Algorithm == DFX_SSAO_ALGORITHM_VBAO)
                    {
                        Visibility += 1.0f - float(__builtin_popcount(OccludedBitfield)) / 32.0f;
                    }
                    else if (A.Algorithm == DFX_SSAO_ALGORITHM_HBAO)
                    {
                        float2 HorizonAngles(+FastACos(MaxCosHorizons.x), -FastACos(MaxCosHorizons.y));
                        Visibility += 0.5f * IntegrateArcUniform(HorizonAngles.x, HorizonAngles.y);
                    }
                    else
                    {
                        float2 HorizonAngles(+FastACos(MaxCosHorizons.x), -FastACos(MaxCosHorizons.y));
                        Visibility += ProjNormalLen * IntegrateArcCosWeighted(HorizonAngles.x, HorizonAngles.y, N, CosNorm);
                    }
                }
                out.at(px, py) = Visibility / 3.0f;
            }
    });
}

// SSAO_ComputeTemporalAccumulation.fx
void ssao_temporal(const Camera& curr, const Camera& prev, const dfx_ssao_attribs& A, const TexF& curr_occ, const TexF& prev_occ,
                   const TexF& prev_hist, const TexF& curr_depth /*reprojected*/, const TexF& prev_depth, const TexF2& motion,
                   TexF& out_occ, TexF& out_hist, int threads)
{
    const int W = curr_occ.w, H = curr_occ.h;
    out_occ.resize(W, H, 1.0f);
    out_hist.resize(W, H, 1.0f);
    const int2 Dim(int(curr.f4ViewportSize.x), int(curr.f4ViewportSize.y));

    auto IsCameraZSimilar = [](float CurrCamZ, float PrevCamZ) { return std::fabs(1.0f - CurrCamZ / PrevCamZ) < 0.01f ? 1.0f : 0.0f; };

    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                float  Depth = curr_depth.load(px, py);
                if (IsBackground(Depth)) continue; // discard

                float2 Motion       = motion.load(px, py) * float2(F3NDC_XYZ_TO_UVD_SCALE.x, F3NDC_XYZ_TO_UVD_SCALE.y);
                float2 PrevLocation = Position - Motion * float2(curr.f4ViewportSize.x, curr.f4ViewportSize.y);

                // ComputeReprojection :105-149
                float        CurrCamZ = DepthToCameraZ(Depth, curr.mProj);
                BilinearInfo b        = GetBilinearSamplingInfoUC(PrevLocation, Dim);
                float        z00      = DepthToCameraZ(prev_depth.load(b.x0, b.y0), prev.mProj);
                float        z10      = DepthToCameraZ(prev_depth.load(b.x1, b.y0), prev.mProj);
                float        z01      = DepthToCameraZ(prev_depth.load(b.x0, b.y1), prev.mProj);
                float        z11      = DepthToCameraZ(prev_depth.load(b.x1, b.y1), prev.mProj);
                float4       Weights(b.w[0], b.w[1], b.w[2], b.w[3]);
                Weights.x *= IsCameraZSimilar(CurrCamZ, z00);
                Weights.y *= IsCameraZSimilar(CurrCamZ, z10);
                Weights.z *= IsCameraZSimilar(CurrCamZ, z01);
                Weights.w *= IsCameraZSimilar(CurrCamZ, z11);
                float TotalWeight = dot(Weights, float4(1, 1, 1, 1));

                float ROcclusion = 1.0f, RHistory = 1.0f;
                bool  IsSuccess  = TotalWeight > 0.01f && !A.ResetAccumulation;
                if (IsSuccess)
                {
                    float4 PrevOcclusion(prev_occ.load(b.x0, b.y0), prev_occ.load(b.x1, b.y0), prev_occ.load(b.x0, b.y1), prev_occ.load(b.x1, b.y1));
                    float4 History(prev_hist.load(b.x0, b.y0), prev_hist.load(b.x1, b.y0), prev_hist.load(b.x0, b.y1), prev_hist.load(b.x1, b.y1));
                    History    = float4(hmin(History.x + 1.0f, 16.0f), hmin(History.y + 1.0f, 16.0f), hmin(History.z + 1.0f, 16.0f), hmin(History.w + 1.0f, 16.0f));
                    ROcclusion = dot(PrevOcclusion, Weights) / TotalWeight;
                    RHistory   = dot(History, Weights) / TotalWeight;
                }

                if (IsSuccess)
                {
                    // ComputePixelStatistic :81-103
                    float M1 = 0.0f, M2 = 0.0f;
                    for (int x = -1; x <= 1; ++x)
                        for (int y = -1; y <= 1; ++y)
                        {
                            int2  L = ClampScreenCoord(int2(px + x, py + y), Dim);
                            float s = curr_occ.load(L);
                            M1 += s;
                            M2 += s * s;
                        }
                    float Mean     = M1 / 9.0f;
                    float Variance = (M2 / 9.0f) - (Mean * Mean);
                    float StdDev   = std::sqrt(hmax(Variance, 0.0f));

                    float AspectRatio   = curr.f4ViewportSize.x * curr.f4ViewportSize.w;
                    float MotionFactor  = saturate(1.025f - length(float2(Motion.x * AspectRatio, Motion.y)) * 128.0f);
                    float VarianceGamma = lerp(0.5f, 2.5f, MotionFactor * MotionFactor);
                    float OcclusionMin  = Mean - VarianceGamma * StdDev;
                    float OcclusionMax  = Mean + VarianceGamma * StdDev;
                    bool  IsInsideRange = OcclusionMin < ROcclusion && ROcclusion < OcclusionMax;
                    RHistory            = IsInsideRange ? RHistory : hmax(1.0f, MotionFactor * RHistory);
                }

                float Alpha        = rcp(RHistory);
                out_occ.at(px, py) = lerp(ROcclusion, curr_occ.load(px, py), Alpha);
                out_hist.at(px, py) = RHistory;
            }
    });
}

void ssao_convolute(const TexF& accumulated, const TexF& depth, MipTex<float>& occ_pyr, MipTex<float>& depth_pyr, int threads)
{
    const int levels = std::min(compute_mip_levels_count(depth.w, depth.h), 5);
    occ_pyr.create(depth.w, depth.h, levels);
    depth_pyr.create(depth.w, depth.h, levels);
    occ_pyr.mip[0].d   = accumulated.d;
    depth_pyr.mip[0].d = depth.d;
    for (int m = 1; m < levels; ++m)
    {
        for (int which = 0; which < 2; ++which)
        {
            const TexF& last = which ? depth_pyr.mip[m - 1] : occ_pyr.mip[m - 1];
            TexF&       dst  = which ? depth_pyr.mip[m] : occ_pyr.mip[m];
            parallel_rows(0, dst.h, threads, [&](int ya, int yb) {
                for (int y = ya; y < yb; ++y)
                    for (int x = 0; x < dst.w; ++x)
                    {
                        float s[9];
                        uint  n = gather_footprint(last, 2 * x, 2 * y, [](float v) { return v; }, s);
                        float r = 0.0f;
                        for (uint i = 0; i < n; ++i) r += s[i];
                        dst.at(x, y) = r / float(n);
                    }
            });
        }
    }
}

void ssao_resample(const Camera& cam, const MipTex<float>& occ_pyr, const MipTex<float>& depth_pyr, const TexF& history,
                   const TexF4& normal, TexF& out, int threads)
{
    const int W = history.w, H = history.h;
    out.resize(W, H);
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                float  Depth              = depth_pyr.mip[0].load(px, py);
                float  History            = history.load(px, py);
                float  AccumulationFactor = (History - 1.0f) / 4.0f;
                if (IsBackground(Depth) || AccumulationFactor >= 1.0f)
                {
                    out.at(px, py) = occ_pyr.mip[0].load(px, py);
                    continue;
                }
                int    MipLevel   = int(4.0f * (1.0f - saturate(AccumulationFactor)));
                float3 PositionSS(Position * float2(cam.f4ViewportSize.z, cam.f4ViewportSize.w), Depth);
                float3 PositionVS = ScreenXYDepthToViewSpace(PositionSS, cam.mProj);
                float3 NormalVS   = mul_dir(normal.load(px, py).xyz(), cam.mView);
                float  PlaneNormalFactor = 10.0f / (1.0f + DepthToCameraZ(Depth, cam.mProj));

                float OcclusionSum = 0.0f, WeightSum = 0.0f;
                // Note: the shader can index a mip the texture does not have when min(mips,5) < 5; sizes used here always have 5.
                MipLevel = std::min(MipLevel, occ_pyr.levels() - 1);
                while (MipLevel >= 0 && WeightSum < 0.995f)
                {
                    float  inv = rcp(float(1u << uint(MipLevel)));
                    float2 MipResolution = float2(cam.f4ViewportSize.x, cam.f4ViewportSize.y) * inv;
                    float2 MipLocation   = Position * inv;
                    int2   MipLocationi(int(MipLocation.x - 0.5f), int(MipLocation.y - 0.5f));
                    float  x = frac(MipLocation.x + 0.5f);
                    float  y = frac(MipLocation.y + 0.5f);
                    float  Weight[4] = {(1.0f - x) * (1.0f - y), x * (1.0f - y), (1.0f - x) * y, x * y};
                    OcclusionSum = 0.0f;
                    WeightSum    = 0.0f;
                    for (int SampleIdx = 0; SampleIdx < 4; SampleIdx++)
                    {
                        int2   Location = MipLocationi + int2(SampleIdx & 0x01, SampleIdx >> 1);
                        float2 Texcoord = (float2(float(Location.x), float(Location.y)) + float2(0.5f, 0.5f)) * float2(rcp(MipResolution.x), rcp(MipResolution.y));
                        float  SampledDepth     = sample_linear(depth_pyr.mip[MipLevel], Texcoord, Address::Clamp); // Sam_LinearClamp (…cpp:735)
                        float  SampledOcclusion = sample_point_clamp(occ_pyr.mip[MipLevel], Texcoord);              // Sam_PointClamp  (…cpp:736)
                        float3 SamplePositionVS = ScreenXYDepthToViewSpace(float3(Texcoord, SampledDepth), cam.mProj);
                        float  WeightS = Weight[SampleIdx];
                        float  WeightZ = ComputeGeometryWeight(PositionVS, SamplePositionVS, NormalVS, PlaneNormalFactor);
                        OcclusionSum += SampledOcclusion * WeightS * WeightZ;
                        WeightSum += WeightS * WeightZ;
                    }
                    MipLevel--;
                }
                out.at(px, py) = OcclusionSum / WeightSum;
            }
    });
}

void ssao_spatial(const Camera& cam, const dfx_ssao_attribs& A, const TexF& occlusion, const TexF& history, const TexF& depth,
                  const TexF4& normal, TexF& out, int threads)
{
    static const float3 Poisson[8] = {
        float3(-0.4706069f, -0.4427112f, +0.6461146f), float3(-0.9057375f, +0.3003471f, +0.9542373f),
        float3(-0.3487388f, +0.4037880f, +0.5335386f), float3(+0.1023042f, +0.6439373f, +0.6520134f),
        float3(+0.5699277f, +0.3513750f, +0.6695386f), float3(+0.2939128f, -0.1131226f, +0.3149309f),
        float3(+0.7836658f, -0.4208784f, +0.8895339f), float3(+0.1564120f, -0.8198990f, +0.8346850f)};
    const int W = depth.w, H = depth.h;
    out.resize(W, H);
    const int2 Dim(int(cam.f4ViewportSize.x), int(cam.f4ViewportSize.y));
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                float  History = history.load(px, py);
                float  Depth   = depth.load(px, py);
                float  AccumulationFactor = std::pow(std::fabs((History - 1.0f) / 8.0f), 0.2f);
                if (IsBackground(Depth) || AccumulationFactor >= 1.0f)
                {
                    out.at(px, py) = lerp(1.0f, occlusion.load(px, py), A.AlphaInterpolation);
                    continue;
                }
                float3 PositionSS(Position * float2(cam.f4ViewportSize.z, cam.f4ViewportSize.w), Depth);
                float3 PositionVS = ScreenXYDepthToViewSpace(PositionSS, cam.mProj);
                float3 NormalVS   = mul_dir(normal.load(px, py).xyz(), cam.mView);
                float4 Rotator    = GetRotator(2.0f * M_PI_F * Bayer4x4(uint(px), uint(py), cam.uiFrameIndex));
                float  Radius     = lerp(0.0f, A.SpatialReconstructionRadius, 1.0f - saturate(AccumulationFactor));
                float  PlaneNormalFactor = 10.0f / (1.0f + DepthToCameraZ(Depth, cam.mProj));

                float OcclusionSum = 0.0f, WeightSum = 0.0f;
                for (int i = 0; i < 8; i++)
                {
                    float2 Xi = RotateVector(Rotator, Poisson[i].xy());
                    float2 sp = Position + Radius * Xi;
                    int2   SampleCoord = ClampScreenCoord(int2(ftoi(sp.x), ftoi(sp.y)), Dim);
                    float  SampledDepth     = depth.load(SampleCoord);
                    float  SampledOcclusion = occlusion.load(SampleCoord);
                    float3 SamplePositionSS((float2(float(SampleCoord.x), float(SampleCoord.y)) + float2(0.5f, 0.5f)) * float2(cam.f4ViewportSize.z, cam.f4ViewportSize.w), SampledDepth);
                    float3 SamplePositionVS = ScreenXYDepthToViewSpace(SamplePositionSS, cam.mProj);
                    float  WeightS = ComputeSpatialWeight(Poisson[i].z * Poisson[i].z, 0.9f);
                    float  WeightZ = ComputeGeometryWeight(PositionVS, SamplePositionVS, NormalVS, PlaneNormalFactor);
                    OcclusionSum += WeightS * WeightZ * SampledOcclusion;
                    WeightSum += WeightS * WeightZ;
                }
                float Occlusion = WeightSum > 0.0f ? OcclusionSum / WeightSum : occlusion.load(px, py);
                out.at(px, py)  = lerp(1.0f, Occlusion, A.AlphaInterpolation);
            }
    });
}

} // namespace orc
