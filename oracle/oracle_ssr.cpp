// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h). Pinned against the reference's own shaders (oracle/refshader, tests/test_reference_shaders.py).
// SSR passes S1, S2, S4-S7 restated from Shaders/PostProcess/ScreenSpaceReflection/private/*.fx (file:line per function).
#include "oracle.h"
#include <atomic>

namespace orc
{

static inline bool IsBackground(float Depth) { return g_reversed_depth ? Depth < 1e-6f : Depth >= (1.0f - 1e-6f); } // SSR_Common.fxh:48-55
static inline bool IsReflectionSample(float R, float D, float Thr) { return R <= Thr && !IsBackground(D); } // :57-60
static inline bool IsMirrorReflection(float R) { return R < 0.01f; }                                     // :62-65

// ---------------------------------------------------------------------------------------------------------------------
// S1  SSR_ComputeHierarchicalDepthBuffer.fx:30-73
void ssr_hiz(const TexF& depth, MipTex<float>& pyr, int threads)
{
    const int levels = std::min(compute_mip_levels_count(depth.w, depth.h), 7); // SSR_DEPTH_HIERARCHY_MAX_MIP + 1
    pyr.create(depth.w, depth.h, levels);
    pyr.mip[0].d = depth.d;
    for (int m = 1; m < levels; ++m)
    {
        const TexF& last = pyr.mip[m - 1];
        TexF&       dst  = pyr.mip[m];
        parallel_rows(0, dst.h, threads, [&](int ya, int yb) {
            for (int y = ya; y < yb; ++y)
                for (int x = 0; x < dst.w; ++x)
                {
                    int   rx = 2 * x, ry = 2 * y;
                    float MinDepth = g_reversed_depth ? 0.0f : 1.0f; // DepthFarPlane; ClosestDepth = max when reversed (SSR_Common.fxh:6-12)
                    auto  upd      = [&](int ox, int oy) {
                        const float d = last.load_clamped(rx + ox, ry + oy);
                        MinDepth      = g_reversed_depth ? hmax(MinDepth, d) : hmin(MinDepth, d);
                    };
                    upd(0, 0);
                    upd(0, 1);
                    upd(1, 0);
                    upd(1, 1);
                    bool wo = (last.w & 1) != 0, ho = (last.h & 1) != 0;
                    if (wo)
                    {
                        upd(2, 0);
                        upd(2, 1);
                    }
                    if (ho)
                    {
                        upd(0, 2);
                        upd(1, 2);
                    }
                    if (wo && ho) upd(2, 2);
                    dst.at(x, y) = MinDepth;
                }
        });
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// S2  SSR_ComputeStencilMaskAndExtractRoughness.fx:13-40 ; host: ScreenSpaceReflection.cpp:904-932
void ssr_mask_roughness(const dfx_ssr_attribs& A, const TexF4& material, const TexF& depth, TexF& roughness, Tex<uint8_t>& mask, int threads)
{
    const int W = depth.w, H = depth.h;
    if (roughness.w != W || roughness.h != H) roughness.resize(W, H, 0.0f); // never cleared afterwards (stale where masked)
    mask.resize(W, H, 0);                                                     // ClearDepthStencil 0.0
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < W; ++x)
            {
                float4 M = material.load(x, y);
                float4 Sel(A.RoughnessChannel == 0u ? 1.0f : 0.0f, A.RoughnessChannel == 1u ? 1.0f : 0.0f,
                           A.RoughnessChannel == 2u ? 1.0f : 0.0f, A.RoughnessChannel == 3u ? 1.0f : 0.0f);
                float Roughness = dot(M, Sel);
                if (!A.IsRoughnessPerceptual) Roughness = std::sqrt(Roughness);
                float Depth = depth.load(x, y);
                if (!IsReflectionSample(Roughness, Depth, A.RoughnessThreshold)) continue; // discard
                roughness.at(x, y) = Roughness;
                mask.at(x, y)      = 1;
            }
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// S3  SSR_ComputeDownsampledStencilMask.fx:13-61 ; host: ScreenSpaceReflection.cpp:934-961 (half resolution only)
void ssr_downsample_mask(const dfx_ssr_attribs& A, const TexF& roughness, const TexF& depth, Tex<uint8_t>& mask_half, int threads)
{
    const int W = depth.w / 2, H = depth.h / 2;
    mask_half.resize(W, H, 0); // ClearDepthStencil 0.0
    const bool IsWidthOdd = (depth.w & 1) != 0, IsHeightOdd = (depth.h & 1) != 0;
    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int y = ya; y < yb; ++y)
            for (int x = 0; x < W; ++x)
            {
                float MinDepth = g_reversed_depth ? 0.0f : 1.0f, MaxRoughness = 0.0f; // DepthFarPlane
                auto  upd      = [&](int ox, int oy) {
                    const float d = depth.load_clamped(2 * x + ox, 2 * y + oy);
                    MinDepth      = g_reversed_depth ? hmax(MinDepth, d) : hmin(MinDepth, d);
                    MaxRoughness  = hmax(MaxRoughness, roughness.load_clamped(2 * x + ox, 2 * y + oy));
                };
                upd(0, 0), upd(1, 0), upd(0, 1), upd(1, 1);
                if (IsWidthOdd) upd(2, 0), upd(2, 1);
                if (IsHeightOdd) upd(0, 2), upd(1, 2);
                if (IsWidthOdd && IsHeightOdd) upd(2, 2);
                if (IsReflectionSample(MaxRoughness, MinDepth, A.RoughnessThreshold)) mask_half.at(x, y) = 1;
            }
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// S4  SSR_ComputeIntersection.fx
// workload statistics of the Hi-Z march (rays traced, loop iterations), for DESIGN.md / the bench report
std::atomic<unsigned long long> g_march_rays{0}, g_march_iterations{0};
// optional per-pixel record of the loop's trip count (0 where no ray is traced), W x H uint16: what a SIMD implementation needs to
// know about the divergence of neighbouring rays (orc_march_iteration_plane)
unsigned short* g_march_iteration_plane = nullptr;
int             g_march_iteration_pitch = 0;
static thread_local unsigned t_last_march_iterations = 0;

namespace
{

inline float LoadDepthHierarchy(const MipTex<float>& hiz, int x, int y, int mip)
{
    if (mip < 0 || mip >= hiz.levels()) return 0.0f; // Load from a non-existent mip returns 0
    return hiz.mip[mip].load(x, y);
}

// :88-136 AdvanceRay
inline bool AdvanceRay(float3 Origin, float3 Direction, float3 InvDirection, float2 CurrentMipPosition, float2 InvCurrentMipResolution,
                       float2 FloorOffset, float2 UVOffset, float SurfaceDepth, float3& Position, float& CurrentT)
{
    float2 XYPlane = floor2(CurrentMipPosition) + FloorOffset;
    XYPlane        = XYPlane * InvCurrentMipResolution + UVOffset;
    float3 BoundaryPlanes(XYPlane, SurfaceDepth);
    float3 T = BoundaryPlanes * InvDirection - Origin * InvDirection;
    T.z      = (g_reversed_depth ? Direction.z < 0.0f : Direction.z > 0.0f) ? T.z : FLT_MAX_F; // :109-113
    float TMin = hmin(hmin(T.x, T.y), T.z);
    bool  AboveSurface = g_reversed_depth ? SurfaceDepth < Position.z : SurfaceDepth > Position.z; // :118-124
    bool  SkippedTile  = asuint(TMin) != asuint(T.z) && AboveSurface;
    CurrentT           = AboveSurface ? TMin : CurrentT;
    Position           = Origin + CurrentT * Direction;
    return SkippedTile;
}

// :139-189 HierarchicalRaymarch
inline float3 HierarchicalRaymarch(const MipTex<float>& hiz, float3 Origin, float3 Direction, float2 ScreenSize, int MostDetailedMip,
                                   uint MaxTraversalIntersections, bool& ValidHit)
{
    float3 InvDirection(Direction.x != 0.0f ? rcp(Direction.x) : FLT_MAX_F, Direction.y != 0.0f ? rcp(Direction.y) : FLT_MAX_F,
                        Direction.z != 0.0f ? rcp(Direction.z) : FLT_MAX_F);
    int    CurrentMip = MostDetailedMip;
    float2 CurrentMipResolution    = ScreenSize * rcp(float(1 << CurrentMip));
    float2 InvCurrentMipResolution(rcp(CurrentMipResolution.x), rcp(CurrentMipResolution.y));

    float2 UVOffset = 0.005f * float(1 << MostDetailedMip) / ScreenSize;
    UVOffset.x      = Direction.x < 0.0f ? -UVOffset.x : UVOffset.x;
    UVOffset.y      = Direction.y < 0.0f ? -UVOffset.y : UVOffset.y;
    float2 FloorOffset(Direction.x < 0.0f ? 0.0f : 1.0f, Direction.y < 0.0f ? 0.0f : 1.0f);

    // InitialAdvanceRay :66-86
    float  CurrentT;
    float3 Position;
    {
        float2 CurrentMipPosition = CurrentMipResolution * Origin.xy();
        float2 XYPlane            = floor2(CurrentMipPosition) + FloorOffset;
        XYPlane                   = XYPlane * InvCurrentMipResolution + UVOffset;
        float2 T                  = XYPlane * InvDirection.xy() - Origin.xy() * InvDirection.xy();
        CurrentT                  = hmin(T.x, T.y);
        Position                  = Origin + CurrentT * Direction;
    }

    uint Idx = 0u;
    while (Idx < MaxTraversalIntersections && CurrentMip >= MostDetailedMip)
    {
        float2 CurrentMipPosition = CurrentMipResolution * Position.xy();
        float  SurfaceDepth       = LoadDepthHierarchy(hiz, ftoi(CurrentMipPosition.x), ftoi(CurrentMipPosition.y), CurrentMip);
        bool   SkippedTile = AdvanceRay(Origin, Direction, InvDirection, CurrentMipPosition, InvCurrentMipResolution, FloorOffset, UVOffset,
                                        SurfaceDepth, Position, CurrentT);
        bool NextMipIsOutOfRange = SkippedTile && (CurrentMip >= 6);
        if (!NextMipIsOutOfRange)
        {
            CurrentMip += SkippedTile ? 1 : -1;
            CurrentMipResolution    = CurrentMipResolution * (SkippedTile ? 0.5f : 2.0f);
            InvCurrentMipResolution = InvCurrentMipResolution * (SkippedTile ? 2.0f : 0.5f);
        }
        ++Idx;
    }
    ValidHit = (Idx <= MaxTraversalIntersections);
    g_march_rays.fetch_add(1, std::memory_order_relaxed);
    g_march_iterations.fetch_add(Idx, std::memory_order_relaxed);
    t_last_march_iterations = Idx;
    return Position;
}

// :192-197
inline float CalculateEdgeVignette(float2 Hit, float2 ScreenSize)
{
    float2 FOV = 0.05f * float2(ScreenSize.y / ScreenSize.x, 1.0f);
    float  bx  = smoothstep(0.0f, FOV.x, Hit.x) * (1.0f - smoothstep(1.0f - FOV.x, 1.0f, Hit.x));
    float  by  = smoothstep(0.0f, FOV.y, Hit.y) * (1.0f - smoothstep(1.0f - FOV.y, 1.0f, Hit.y));
    return bx * by;
}
} // namespace

void ssr_intersect(const Camera& cam, const dfx_ssr_attribs& A, uint flags, const TexF4& color, const TexF4& normal, const TexF& roughness,
                   const Tex<uint8_t>& mask, const TexF2& blue_noise_xy, const MipTex<float>& hiz, const TexF2* motion, TexF4& out_radiance,
                   TexF4& out_raydir_pdf, int threads)
{
    // FEATURE_FLAG_HALF_RESOLUTION: targets and (downsampled) mask are W/2 x H/2; every target pixel traces the ray of ONE of
    // its four full-resolution pixels, chosen by a 4x4 pattern (:283-288)
    const bool HalfRes = (flags & DFX_SSR_FEATURE_FLAG_HALF_RESOLUTION) != 0;
    const int  W = HalfRes ? color.w / 2 : color.w, H = HalfRes ? color.h / 2 : color.h;
    out_radiance.resize(W, H, float4());
    out_raydir_pdf.resize(W, H, float4());
    const bool   PreviousFrame = (flags & DFX_SSR_FEATURE_FLAG_PREVIOUS_FRAME) != 0;
    const float2 ScreenSize(cam.f4ViewportSize.x, cam.f4ViewportSize.y);

    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                if (!mask.load(px, py)) continue; // depth test LESS vs mask (early depth-stencil)

                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                if (HalfRes)
                {
                    uint SampleIdx = ComputeHalfResolutionOffset(uint(px), uint(py));
                    Position       = float2(2.0f * float(px) + float(SampleIdx & 0x01u) + 0.5f, 2.0f * float(py) + float(SampleIdx >> 1u) + 0.5f);
                }
                const int fx = ftoi(Position.x), fy = ftoi(Position.y); // int2(Position)
                float2 ScreenCoordUV = Position * float2(cam.f4ViewportSize.z, cam.f4ViewportSize.w);
                float3 NormalWS  = normal.load(fx, fy).xyz();
                float3 NormalVS  = mul_dir(NormalWS, cam.mView);
                float  Roughness = roughness.load(fx, fy);

                bool   IsMirror        = IsMirrorReflection(Roughness);
                int    MostDetailedMip = IsMirror ? 0 : int(A.MostDetailedMip);
                float2 MipResolution   = ScreenSize * rcp(float(1 << MostDetailedMip));

                float2 mp = ScreenCoordUV * MipResolution;
                float3 RayOriginSS(ScreenCoordUV, LoadDepthHierarchy(hiz, ftoi(mp.x), ftoi(mp.y), MostDetailedMip));
                float3 RayOriginVS = ScreenXYDepthToViewSpace(RayOriginSS, cam.mProj);

                // SampleReflectionVector :254-278
                float4 RayDirectionVS;
                {
                    float3 View = -normalize(RayOriginVS);
                    float  AlphaRoughness = Roughness * Roughness;
                    float3 N = NormalVS;
                    float3 T = normalize(cross(N, std::fabs(N.y) > 0.5f ? float3(1.0f, 0.0f, 0.0f) : float3(0.0f, 1.0f, 0.0f)));
                    float3 B = cross(T, N);
                    float2 Xi = blue_noise_xy.load(px & 127, py & 127);
                    Xi.y      = lerp(Xi.y, 0.0f, A.GGXImportanceSampleBias);
                    // mul(TangentToWorld, View) with rows T,B,N
                    float3 ViewDirTS(dot(T, View), dot(B, View), dot(N, View));
                    float3 MicroNormalTS = SmithGGXSampleVisibleNormalSC(ViewDirTS, AlphaRoughness, AlphaRoughness, Xi.x, Xi.y);
                    float3 SampleDirTS   = reflect(-ViewDirTS, MicroNormalTS);
                    float  NdotV = ViewDirTS.z, NdotH = MicroNormalTS.z;
                    float  D   = NormalDistribution_GGX(NdotH, AlphaRoughness);
                    float  G1  = SmithGGXMasking(NdotV, AlphaRoughness);
                    float  PDF = G1 * D / (4.0f * NdotV + FLT_EPS_F);
                    // mul(SampleDirTS, TangentToWorld)
                    float3 dirVS = SampleDirTS.x * T + SampleDirTS.y * B + SampleDirTS.z * N;
                    RayDirectionVS = float4(dirVS, PDF);
                }
                float3 RayDirectionSS = ProjectDirection(RayOriginVS, RayDirectionVS.xyz(), RayOriginSS, cam.mProj);
                float3 RayDirectionWS = mul_dir(RayDirectionVS.xyz(), cam.mViewInv);

                bool   ValidHit = false;
                float3 SurfaceHitSS = HierarchicalRaymarch(hiz, RayOriginSS, RayDirectionSS, ScreenSize, MostDetailedMip, A.MaxTraversalIntersections, ValidHit);
                if (g_march_iteration_plane) g_march_iteration_plane[size_t(py) * g_march_iteration_pitch + px] = (unsigned short)std::min(t_last_march_iterations, 65535u);
                float3 SurfaceHitVS = ScreenXYDepthToViewSpace(SurfaceHitSS, cam.mProj);

                float2 HitPrev = SurfaceHitSS.xy();
                if (PreviousFrame && motion)
                {
                    float2 hp = ScreenSize * SurfaceHitSS.xy();
                    float2 Motion = motion->load(ftoi(hp.x), ftoi(hp.y)) * float2(F3NDC_XYZ_TO_UVD_SCALE.x, F3NDC_XYZ_TO_UVD_SCALE.y);
                    HitPrev       = SurfaceHitSS.xy() - Motion;
                }

                // ValidateHit :201-242
                float Confidence = 0.0f;
                if (ValidHit)
                {
                    float3 Hit = SurfaceHitSS;
                    do
                    {
                        if (Hit.x < 0.0f || Hit.y < 0.0f || Hit.x > 1.0f || Hit.y > 1.0f) break;
                        float2 ManhattanDist(std::fabs(Hit.x - ScreenCoordUV.x), std::fabs(Hit.y - ScreenCoordUV.y));
                        if (ManhattanDist.x < (2.0f / ScreenSize.x) && ManhattanDist.y < (2.0f / ScreenSize.y)) break;
                        float2 tc = ScreenSize * Hit.xy();
                        int2   TexelCoords(ftoi(tc.x), ftoi(tc.y));
                        float  SurfaceDepth = LoadDepthHierarchy(hiz, TexelCoords.x, TexelCoords.y, 0);
                        if (IsBackground(SurfaceDepth)) break;
                        float3 HitNormalWS = normal.load(TexelCoords).xyz();
                        if (dot(HitNormalWS, RayDirectionWS) > 0.0f) break;
                        float3 SurfaceVS = ScreenXYDepthToViewSpace(float3(Hit.xy(), SurfaceDepth), cam.mProj);
                        float3 HitVS     = ScreenXYDepthToViewSpace(Hit, cam.mProj);
                        float  Distance  = distance(SurfaceVS, HitVS);
                        float  Vignette  = PreviousFrame ? hmin(CalculateEdgeVignette(HitPrev, ScreenSize), CalculateEdgeVignette(Hit.xy(), ScreenSize))
                                                         : CalculateEdgeVignette(Hit.xy(), ScreenSize);
                        float  c = 1.0f - smoothstep(0.0f, A.DepthBufferThickness, Distance * rcp(SurfaceVS.z + FLT_EPS_F));
                        c *= c;
                        Confidence = Vignette * c;
                    } while (false);
                }

                float3 ReflectionRadiance(0.0f, 0.0f, 0.0f);
                if (Confidence > 0.0f)
                {
                    float2 rc = ScreenSize * (PreviousFrame ? HitPrev : SurfaceHitSS.xy());
                    ReflectionRadiance = color.load(ftoi(rc.x), ftoi(rc.y)).xyz();
                }
                out_radiance.at(px, py)   = float4(ReflectionRadiance, Confidence);
                out_raydir_pdf.at(px, py) = float4(RayDirectionWS * distance(SurfaceHitVS, RayOriginVS), RayDirectionVS.w);
            }
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// S5  SSR_ComputeSpatialReconstruction.fx:114-172
void ssr_spatial(const Camera& cam, const dfx_ssr_attribs& A, const TexF& roughness, const Tex<uint8_t>& mask, const TexF4& normal,
                 const TexF& depth, const TexF4& raydir_pdf, const TexF4& radiance, TexF4& out_radiance, TexF& out_variance, TexF& out_depth,
                 int threads, bool half_res)
{
    static const float3 Poisson[8] = {
        float3(-0.4706069f, -0.4427112f, +0.6461146f), float3(-0.9057375f, +0.3003471f, +0.9542373f),
        float3(-0.3487388f, +0.4037880f, +0.5335386f), float3(+0.1023042f, +0.6439373f, +0.6520134f),
        float3(+0.5699277f, +0.3513750f, +0.6695386f), float3(+0.2939128f, -0.1131226f, +0.3149309f),
        float3(+0.7836658f, -0.4208784f, +0.8895339f), float3(+0.1564120f, -0.8198990f, +0.8346850f)};
    const int W = depth.w, H = depth.h;
    if (out_radiance.w != W || out_radiance.h != H) out_radiance.resize(W, H, float4());
    if (out_variance.w != W || out_variance.h != H) out_variance.resize(W, H, 0.0f);
    if (out_depth.w != W || out_depth.h != H) out_depth.resize(W, H, 0.0f);
    const int2   Dim(int(cam.f4ViewportSize.x), int(cam.f4ViewportSize.y));
    const float3 CamPos = cam.f4Position.xyz();

    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                if (!mask.load(px, py)) continue;
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);
                float2 ScreenCoordUV = Position * float2(cam.f4ViewportSize.z, cam.f4ViewportSize.w);
                float3 PositionWS = InvProjectPosition(float3(ScreenCoordUV, depth.load(px, py)), cam.mViewProjInv);
                float3 NormalWS   = normal.load(px, py).xyz();
                float3 ViewWS     = normalize(CamPos - PositionWS);
                float  NdotV      = saturate(dot(NormalWS, ViewWS));

                float  Roughness       = roughness.load(px, py);
                float  RoughnessFactor = saturate(5.0f * Roughness);
                float  Radius          = lerp(0.0f, A.SpatialReconstructionRadius, RoughnessFactor);
                float4 Rotator         = GetRotator(2.0f * M_PI_F * Bayer4x4(uint(px), uint(py), cam.uiFrameIndex));

                float4 ColorSum;
                float  WeightSum = 0.0f, Variance = 0.0f, Mean = 0.0f;
                float  NearestSurfaceHitDistance = 0.0f;

                for (int i = 0; i < 8; i++)
                {
                    float2 Xi = RotateVector(Rotator, Poisson[i].xy());
                    float2 sp = Position + Radius * Xi;
                    int2   SampleCoord = ClampScreenCoord(int2(ftoi(sp.x), ftoi(sp.y)), Dim);
                    if (half_res) // the intersect targets are W/2 x H/2 (:153-157)
                    {
                        float2 hp   = (float2(float(px), float(py)) + Radius * Xi) * 0.5f + float2(0.5f, 0.5f);
                        SampleCoord = ClampScreenCoord(int2(ftoi(hp.x), ftoi(hp.y)), int2(int(0.5f * cam.f4ViewportSize.x), int(0.5f * cam.f4ViewportSize.y)));
                    }
                    float  WeightS = ComputeSpatialWeight(Poisson[i].z * Poisson[i].z, 0.9f);

                    // ComputeWeightRayLength :60-86
                    float2 WeightLength;
                    {
                        float4 RayDirectionPDF = raydir_pdf.load(SampleCoord);
                        float  RayLength       = length(RayDirectionPDF.xyz());
                        if (RayLength < 1e-6f)
                        {
                            WeightLength = float2(1e-6f, 1e-6f);
                        }
                        else
                        {
                            float3 L   = RayDirectionPDF.xyz() / RayLength;
                            float  PDF = RayDirectionPDF.w;
                            float  AlphaRoughness = Roughness * Roughness;
                            float3 Hv = normalize(L + ViewWS);
                            float  NdotH = saturate(dot(NormalWS, Hv));
                            float  NdotL = saturate(dot(NormalWS, L));
                            float  Vis = SmithGGXVisibilityCorrelated(NdotL, NdotV, AlphaRoughness);
                            float  D   = NormalDistribution_GGX(NdotH, AlphaRoughness);
                            float  LocalBRDF = Vis * D * NdotL;
                            LocalBRDF *= WeightS;
                            WeightLength = float2(hmax(LocalBRDF / hmax(PDF, 1e-5f), 1e-6f), RayLength);
                        }
                    }
                    float4 SampleColor = radiance.load(SampleCoord);
                    // ComputeWeightedVariance :90-100
                    {
                        float Weight = WeightLength.x;
                        ColorSum += Weight * SampleColor;
                        WeightSum += Weight;
                        float Value    = Luminance(SampleColor.xyz());
                        float PrevMean = Mean;
                        Mean += Weight * rcp(WeightSum) * (Value - PrevMean);
                        Variance += Weight * (Value - PrevMean) * (Value - Mean);
                    }
                    if (WeightLength.x > 1.0e-6f) NearestSurfaceHitDistance = hmax(WeightLength.y, NearestSurfaceHitDistance);
                }

                out_radiance.at(px, py) = ColorSum / hmax(WeightSum, 1e-6f);
                out_variance.at(px, py) = Variance / hmax(WeightSum, 1e-6f);
                // ComputeResolvedDepth :102-106
                float CameraSurfaceDistance = distance(CamPos, PositionWS);
                out_depth.at(px, py)        = CameraZToDepth(CameraSurfaceDistance + NearestSurfaceHitDistance, cam.mProj);
            }
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// S6  SSR_ComputeTemporalAccumulation.fx
void ssr_temporal(const Camera& curr, const Camera& prev, const dfx_ssr_attribs& A, const Tex<uint8_t>& mask, const TexF2& motion,
                  const TexF& hit_depth, const TexF& curr_depth /*reprojected*/, const TexF4& curr_radiance, const TexF& curr_variance,
                  const TexF& prev_depth, const TexF4& prev_radiance, const TexF& prev_variance, TexF4& out_radiance, TexF& out_variance,
                  int threads)
{
    const int W = curr_depth.w, H = curr_depth.h;
    if (out_radiance.w != W || out_radiance.h != H) out_radiance.resize(W, H, float4());
    if (out_variance.w != W || out_variance.h != H) out_variance.resize(W, H, 0.0f);
    const int2   Dim(int(curr.f4ViewportSize.x), int(curr.f4ViewportSize.y));
    const int2   DepthDim(curr_depth.w, curr_depth.h);
    const float2 Viewport(curr.f4ViewportSize.x, curr.f4ViewportSize.y);
    const float2 InvViewport(curr.f4ViewportSize.z, curr.f4ViewportSize.w);
    const float2 UVD(F3NDC_XYZ_TO_UVD_SCALE.x, F3NDC_XYZ_TO_UVD_SCALE.y);

    auto ComputeDisocclusion = [](float CurrCameraZ, float PrevCameraZ) {
        CurrCameraZ = std::fabs(CurrCameraZ);
        PrevCameraZ = std::fabs(PrevCameraZ);
        return std::exp(-std::fabs(CurrCameraZ - PrevCameraZ) / hmax(hmax(CurrCameraZ, PrevCameraZ), 1e-6f));
    };
    auto SamplePrevRadianceLinear = [&](float2 PixelCoord) { return sample_linear(prev_radiance, PixelCoord * InvViewport, Address::Clamp); };
    auto SamplePrevVarianceLinear = [&](float2 PixelCoord) { return sample_linear(prev_variance, PixelCoord * InvViewport, Address::Clamp); };

    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                if (!mask.load(px, py)) continue;
                float2 Position(float(px) + 0.5f, float(py) + 0.5f);

                // ComputePixelStatistic :119-145
                float4 M1, M2;
                for (int x = -1; x <= 1; x++)
                    for (int y = -1; y <= 1; y++)
                    {
                        int2   L = ClampScreenCoord(int2(px + x, py + y), Dim);
                        float4 s = curr_radiance.load(L);
                        M1 += s;
                        M2 += s * s;
                    }
                float4 Mean     = M1 / 9.0f;
                float4 Variance = (M2 / 9.0f) - (Mean * Mean);
                float4 StdDev   = sqrt4(max4(Variance, 0.0f));

                float  Depth    = curr_depth.load(px, py);
                float  HitDepth = hit_depth.load(px, py);
                float2 Motion   = motion.load(px, py) * UVD;

                float2 PrevIncidentPoint = Position - Motion * Viewport;
                // ComputeReflectionHitPosition :102-108
                float2 PrevReflectionHit;
                {
                    float2 Texcoord    = (float2(float(px), float(py)) + float2(0.5f, 0.5f)) * InvViewport + UVD * curr.f2Jitter;
                    float3 PositionWS  = InvProjectPosition(float3(Texcoord, HitDepth), curr.mViewProjInv);
                    float3 PrevCoordUV = ProjectPosition(PositionWS, prev.mViewProj);
                    PrevReflectionHit  = (PrevCoordUV.xy() - UVD * prev.f2Jitter) * Viewport;
                }
                float4 PrevColorIncidentPoint = SamplePrevRadianceLinear(PrevIncidentPoint);
                float4 PrevColorReflectionHit = SamplePrevRadianceLinear(PrevReflectionHit);
                float  dI = std::fabs(Luminance(PrevColorIncidentPoint.xyz()) - Luminance(Mean.xyz()));
                float  dR = std::fabs(Luminance(PrevColorReflectionHit.xyz()) - Luminance(Mean.xyz()));
                float2 PrevPos = dI < dR ? PrevIncidentPoint : PrevReflectionHit;

                // ComputeReprojection :147-221
                float  CurrCamZ = DepthToCameraZ(Depth, curr.mProj);
                float4 RColor;
                float2 RPrevCoord;
                bool   RSuccess;
                {
                    float PrevCamZ = DepthToCameraZ(prev_depth.load(ftoi(PrevPos.x), ftoi(PrevPos.y)), prev.mProj);
                    RPrevCoord     = PrevPos;
                    RColor         = SamplePrevRadianceLinear(RPrevCoord);
                    RSuccess       = ComputeDisocclusion(CurrCamZ, PrevCamZ) > 0.9f;
                }
                if (!RSuccess)
                {
                    float        BestW[4] = {0, 0, 0, 0};
                    BilinearInfo BestB{0, 0, 0, 0, {0, 0, 0, 0}};
                    float        BestTotalWeight = 0.0f;
                    for (int y = -1; y <= 1; y++)
                    {
                        for (int x = -1; x <= 1; x++)
                        {
                            float2       Location = PrevPos + float2(float(x), float(y));
                            BilinearInfo b = GetBilinearSamplingInfoUC(Location, DepthDim);
                            float z00 = DepthToCameraZ(prev_depth.load(b.x0, b.y0), prev.mProj);
                            float z10 = DepthToCameraZ(prev_depth.load(b.x1, b.y0), prev.mProj);
                            float z01 = DepthToCameraZ(prev_depth.load(b.x0, b.y1), prev.mProj);
                            float z11 = DepthToCameraZ(prev_depth.load(b.x1, b.y1), prev.mProj);
                            float w[4] = {b.w[0], b.w[1], b.w[2], b.w[3]};
                            w[0] *= ComputeDisocclusion(CurrCamZ, z00) > 0.45f ? 1.0f : 0.0f;
                            w[1] *= ComputeDisocclusion(CurrCamZ, z10) > 0.45f ? 1.0f : 0.0f;
                            w[2] *= ComputeDisocclusion(CurrCamZ, z01) > 0.45f ? 1.0f : 0.0f;
                            w[3] *= ComputeDisocclusion(CurrCamZ, z11) > 0.45f ? 1.0f : 0.0f;
                            float TotalWeight = dot(float4(w[0], w[1], w[2], w[3]), float4(1, 1, 1, 1));
                            if (TotalWeight > BestTotalWeight)
                            {
                                BestTotalWeight = TotalWeight;
                                for (int k = 0; k < 4; ++k) BestW[k] = w[k];
                                BestB      = b;
                                RPrevCoord = Location;
                                if (BestTotalWeight > 0.9f) break;
                            }
                        }
                        if (BestTotalWeight > 0.9f) break;
                    }
                    RSuccess = BestTotalWeight > 0.1f;
                    if (RSuccess)
                    {
                        RColor = (prev_radiance.load(BestB.x0, BestB.y0) * BestW[0] + prev_radiance.load(BestB.x1, BestB.y0) * BestW[1] +
                                  prev_radiance.load(BestB.x0, BestB.y1) * BestW[2] + prev_radiance.load(BestB.x1, BestB.y1) * BestW[3]) /
                                 BestTotalWeight;
                    }
                }
                RSuccess = RSuccess && IsInsideScreen(RPrevCoord, Viewport);

                if (RSuccess)
                {
                    float4 ColorMin = Mean - 2.5f * StdDev;
                    float4 ColorMax = Mean + 2.5f * StdDev;
                    float4 PrevRadiance = clamp4(RColor, ColorMin, ColorMax);
                    float  PrevVariance = SamplePrevVarianceLinear(RPrevCoord);
                    out_radiance.at(px, py) = lerp(curr_radiance.load(px, py), PrevRadiance, A.TemporalRadianceStabilityFactor);
                    out_variance.at(px, py) = lerp(curr_variance.load(px, py), PrevVariance, A.TemporalVarianceStabilityFactor);
                }
                else
                {
                    out_radiance.at(px, py) = curr_radiance.load(px, py);
                    out_variance.at(px, py) = 1.0f;
                }
            }
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// S7  SSR_ComputeBilateralCleanup.fx:49-97.  ddx/ddy: 2x2 quad finite differences (fine == coarse for the value used
// by pixel (x,y): v(x|1) - v(x&~1), v(y|1) - v(y&~1); helper lanes evaluate CameraZ regardless of the mask).
void ssr_bilateral(const Camera& cam, const dfx_ssr_attribs& A, const Tex<uint8_t>& mask, const TexF& depth, const TexF4& normal,
                   const TexF& roughness, const TexF4& radiance, const TexF& variance, TexF4& out, int threads)
{
    const int W = depth.w, H = depth.h;
    out.resize(W, H, float4()); // ClearRenderTarget 0
    const int2 Dim(int(cam.f4ViewportSize.x), int(cam.f4ViewportSize.y));
    // quad partner of ddx / ddy (:59): a partner beyond the right / bottom edge of an odd-sized target is a helper lane whose
    // Load returns 0 - confirmed against the shader itself (tests/test_reference_shaders.py::test_bilateral_quad_derivatives_everywhere)
    auto CamZ = [&](int x, int y) { return DepthToCameraZ(depth.load(x, y), cam.mProj); };

    parallel_rows(0, H, threads, [&](int ya, int yb) {
        for (int py = ya; py < yb; ++py)
            for (int px = 0; px < W; ++px)
            {
                if (!mask.load(px, py)) continue;
                float  Roughness = roughness.load(px, py);
                float  Variance  = variance.load(px, py);
                float3 NormalWS  = normal.load(px, py).xyz();
                float  CameraZ   = DepthToCameraZ(depth.load(px, py), cam.mProj);
                float2 GradCamZ(CamZ(px | 1, py) - CamZ(px & ~1, py), CamZ(px, py | 1) - CamZ(px, py & ~1));

                float RoughnessTarget = saturate(8.0f * Roughness);
                float Radius = lerp(0.0f, Variance > 0.001f ? 2.0f : 0.0f, RoughnessTarget);
                float Sigma  = A.BilateralCleanupSpatialSigmaFactor;
                int   EffectiveRadius = int(hmin(2.0f * Sigma, Radius));
                float4 RadianceResult = radiance.load(px, py);

                if (Variance > 0.00005f && EffectiveRadius > 0)
                {
                    float4 ColorSum;
                    float  WeightSum = 0.0f;
                    for (int x = -EffectiveRadius; x <= EffectiveRadius; x++)
                        for (int y = -EffectiveRadius; y <= EffectiveRadius; y++)
                        {
                            int2   L = ClampScreenCoord(int2(px + x, py + y), Dim);
                            float  SampledDepth     = depth.load(L);
                            float  SampledRoughness = roughness.load(L);
                            float4 SampledRadiance  = radiance.load(L);
                            float3 SampledNormalWS  = normal.load(L).xyz();
                            if (IsReflectionSample(SampledRoughness, SampledDepth, A.RoughnessThreshold))
                            {
                                float  SampledCameraZ = DepthToCameraZ(SampledDepth, cam.mProj);
                                float2 xy{float(x), float(y)};
                                float  WeightS = std::exp(-0.5f * dot(xy, xy) / (Sigma * Sigma));
                                float  WeightZ = std::exp(-std::fabs(CameraZ - SampledCameraZ) / (1.0f * (std::fabs(dot(xy, GradCamZ)) + 1e-6f)));
                                float  WeightN = std::pow(hmax(0.0f, dot(NormalWS, SampledNormalWS)), 128.0f);
                                float  Weight  = WeightS * WeightN * WeightZ;
                                WeightSum += Weight;
                                ColorSum += Weight * SampledRadiance;
                            }
                        }
                    RadianceResult = ColorSum / hmax(WeightSum, 1.0e-6f);
                }
                out.at(px, py) = float4(RadianceResult.xyz(), RadianceResult.w * A.AlphaInterpolation);
            }
    });
}

} // namespace orc
