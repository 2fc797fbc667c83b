// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h). Pinned against the reference's own shaders (oracle/refshader, tests/test_reference_shaders.py).
//
// Restatement of the shared shader helpers:
//   Shaders/Common/public/PostFX_Common.fxh, ShaderUtilities.fxh, SRGBUtilities.fxh, PBR_Common.fxh (GGX subset)
// and of the DiligentCore HLSL macros the tree never defines, fixed to the D3D/Vulkan flavour:
//   NormalizedDeviceXYToTexUV(xy) = 0.5 + (0.5,-0.5)*xy       TexUVToNormalizedDeviceXY(uv) = (uv-0.5)*(2,-2)
//   NormalizedDeviceZToDepth(z) = z   DepthToNormalizedDeviceZ(d) = d   F3NDC_XYZ_TO_UVD_SCALE = (0.5,-0.5,1)
//   NDC_MIN_Z = 0    MATRIX_ELEMENT(M,r,c) = M[r][c]    MatrixFromRows(a,b,c) = rows a,b,c
#pragma once
#include "oracle_math.h"
#include "oracle_tex.h"
#include "../include/dfx_b200.h"

namespace orc
{

// PostFXContext::FEATURE_FLAG_REVERSED_DEPTH (PostFXContext.hpp:54): the reference compiles the *_OPTION_INVERTED_DEPTH variants
// of its shaders; the oracle switches at run time (orc_set_reversed_depth). Defined in oracle_capi.cpp.
extern bool g_reversed_depth;


constexpr float M_PI_F       = 3.14159265358979f;   // PostFX_Common.fxh:6  M_PI
constexpr float M_HALF_PI_F  = 1.57079632679490f;   // PostFX_Common.fxh:7
constexpr float PBR_PI       = 3.141592653589793f;  // PBR_Common.fxh:5
constexpr float FLT_EPS_F    = 5.960464478e-8f;     // PostFX_Common.fxh:11
constexpr float FLT_MAX_F    = 3.402823466e+38f;    // PostFX_Common.fxh:12

struct Camera
{
    float4   f4Position;
    float4   f4ViewportSize;
    uint     uiFrameIndex = 0;
    float2   f2Jitter;
    float    fFocusDistance = 10.0f, fFStop = 5.6f, fFocalLength = 50.0f, fSensorWidth = 36.0f; // lens (BasicStructures.fxh:112-121), used by DepthOfField
    float4x4 mView, mProj, mViewProj, mViewInv, mProjInv, mViewProjInv;
};

inline float4x4 to_mat(const dfx_float4x4& s)
{
    float4x4 r;
    std::memcpy(r.m, s.m, sizeof(r.m));
    return r;
}
inline Camera to_camera(const dfx_camera_attribs& c)
{
    Camera r;
    r.f4Position     = {c.f4Position[0], c.f4Position[1], c.f4Position[2], c.f4Position[3]};
    r.f4ViewportSize = {c.f4ViewportSize[0], c.f4ViewportSize[1], c.f4ViewportSize[2], c.f4ViewportSize[3]};
    r.uiFrameIndex   = c.uiFrameIndex;
    r.f2Jitter       = {c.f2Jitter[0], c.f2Jitter[1]};
    r.fFocusDistance = c.fFocusDistance, r.fFStop = c.fFStop, r.fFocalLength = c.fFocalLength, r.fSensorWidth = c.fSensorWidth;
    r.mView          = to_mat(c.mView);
    r.mProj          = to_mat(c.mProj);
    r.mViewProj      = to_mat(c.mViewProj);
    r.mViewInv       = to_mat(c.mViewInv);
    r.mProjInv       = to_mat(c.mProjInv);
    r.mViewProjInv   = to_mat(c.mViewProjInv);
    return r;
}

// ---- DiligentCore macros, D3D/Vulkan flavour ----
inline float2 NormalizedDeviceXYToTexUV(float2 xy) { return float2(0.5f, 0.5f) + float2(0.5f, -0.5f) * xy; }
inline float2 TexUVToNormalizedDeviceXY(float2 uv) { return (uv - float2(0.5f, 0.5f)) * float2(2.0f, -2.0f); }
inline float  NormalizedDeviceZToDepth(float z) { return z; }
inline float  DepthToNormalizedDeviceZ(float d) { return d; }
const float3  F3NDC_XYZ_TO_UVD_SCALE = float3(0.5f, -0.5f, 1.0f);

// ---- ShaderUtilities.fxh ----
// :5-14
inline float CameraZToNormalizedDeviceZ(float CameraZ, const float4x4& mProj)
{
    float m22 = mProj.m[2][2], m32 = mProj.m[3][2], m23 = mProj.m[2][3], m33 = mProj.m[3][3];
    return (m22 * CameraZ + m32) / (m23 * CameraZ + m33);
}
// :16-22
inline float CameraZToDepth(float CameraZ, const float4x4& mProj) { return NormalizedDeviceZToDepth(CameraZToNormalizedDeviceZ(CameraZ, mProj)); }
// :24-31
inline float NormalizedDeviceZToCameraZ(float NdcZ, const float4x4& mProj)
{
    float m22 = mProj.m[2][2], m32 = mProj.m[3][2], m23 = mProj.m[2][3], m33 = mProj.m[3][3];
    return (m32 - NdcZ * m33) / (NdcZ * m23 - m22);
}
// :33-39
inline float DepthToCameraZ(float fDepth, const float4x4& mProj) { return NormalizedDeviceZToCameraZ(DepthToNormalizedDeviceZ(fDepth), mProj); }

// :126-142 GetBilinearSamplingInfoUC. fetch = (x0, y0, x1, y1), weights = (w00, w10, w01, w11)
struct BilinearInfo
{
    int   x0, y0, x1, y1;
    float w[4];
};
inline BilinearInfo GetBilinearSamplingInfoUC(float2 Location, int2 Dimensions)
{
    Location        = Location - float2(0.5f, 0.5f);
    float2 Loc00    = floor2(Location);
    BilinearInfo b;
    b.x0 = int(Loc00.x);
    b.y0 = int(Loc00.y);
    b.x1 = b.x0 + 1;
    b.y1 = b.y0 + 1;
    b.x0 = clampi(b.x0, 0, Dimensions.x - 1);
    b.y0 = clampi(b.y0, 0, Dimensions.y - 1);
    b.x1 = clampi(b.x1, 0, Dimensions.x - 1);
    b.y1 = clampi(b.y1, 0, Dimensions.y - 1);
    float x = Location.x - Loc00.x;
    float y = Location.y - Loc00.y;
    b.w[0] = (1.0f - x) * (1.0f - y);
    b.w[1] = x * (1.0f - y);
    b.w[2] = (1.0f - x) * y;
    b.w[3] = x * y;
    return b;
}

// ---- PostFX_Common.fxh ----
// :20-25
inline uint PCGHash(uint Seed)
{
    uint State = Seed * 747796405u + 2891336453u;
    uint Word  = ((State >> ((State >> 28u) + 4u)) ^ State) * 277803737u;
    return (Word >> 22u) ^ Word;
}
// :40-43
inline float Luminance(float3 Color) { return dot(Color, float3(0.299f, 0.587f, 0.114f)); }
// :57-65
inline float Bayer4x4(uint px, uint py, uint FrameIndex)
{
    uint wx = px & 3u, wy = py & 3u;
    uint A  = 2068378560u * (1u - (wx >> 1u)) + 1500172770u * (wx >> 1u);
    uint B  = (wy + ((wx & 1u) << 2u)) << 2u;
    uint Bayer = ((A >> B) + FrameIndex) & 0xFu;
    return float(Bayer) / 16.0f;
}
// :67-73 GetRotator -> (cos, sin, -sin, cos)
inline float4 GetRotator(float Angle)
{
    float s = std::sin(Angle), c = std::cos(Angle);
    return float4(c, s, -s, c);
}
// :80-83
inline float2 RotateVector(float4 Rotator, float2 Vec) { return Vec.x * float2(Rotator.x, Rotator.z) + Vec.y * float2(Rotator.y, Rotator.w); }
// :85-92
inline float3 ProjectPosition(float3 Origin, const float4x4& Transform)
{
    float4 P  = mul(float4(Origin, 1.0f), Transform);
    float3 p3 = P.xyz() / P.w;
    float2 uv = NormalizedDeviceXYToTexUV(p3.xy());
    return float3(uv, NormalizedDeviceZToDepth(p3.z));
}
// :94-97
inline float3 ProjectDirection(float3 Origin, float3 Direction, float3 OriginSS, const float4x4& Mat) { return ProjectPosition(Origin + Direction, Mat) - OriginSS; }
// :99-105
inline float3 InvProjectPosition(float3 Coord, const float4x4& Transform)
{
    float2 xy = TexUVToNormalizedDeviceXY(Coord.xy());
    float  z  = DepthToNormalizedDeviceZ(Coord.z);
    float4 P  = mul(float4(xy.x, xy.y, z, 1.0f), Transform);
    return P.xyz() / P.w;
}
// :107-111
inline float3 ScreenXYDepthToViewSpace(float3 Coord, const float4x4& Transform)
{
    float2 ndcxy = TexUVToNormalizedDeviceXY(Coord.xy());
    float  ndcz  = DepthToCameraZ(Coord.z, Transform);
    return float3(ndcz * ndcxy.x / Transform.m[0][0], ndcz * ndcxy.y / Transform.m[1][1], ndcz);
}
// :113-127
inline bool IsInsideScreen(float2 PixelCoord, float2 Dimension)
{
    return PixelCoord.x >= 0.0f && PixelCoord.y >= 0.0f && PixelCoord.x < Dimension.x && PixelCoord.y < Dimension.y;
}
// :129-132
// PostFX_Common.fxh:45-55
inline uint ComputeHalfResolutionOffset(uint x, uint y)
{
    const uint PackedOffsets = 1320229860u; // 4x4 matrix of 2-bit offsets: 0 1 2 3 / 3 2 1 0 / 1 0 3 2 / 2 3 0 1
    const uint Idx           = ((x & 0x3u) << 3u) + ((y & 0x3u) << 1u);
    return (PackedOffsets >> Idx) & 0x3u;
}
inline int2 ClampScreenCoord(int2 PixelCoord, int2 Dimension) { return int2(clampi(PixelCoord.x, 0, Dimension.x - 1), clampi(PixelCoord.y, 0, Dimension.y - 1)); }
// :134-137
inline float ComputeSpatialWeight(float Distance, float Sigma) { return std::exp(-(Distance) / (2.0f * Sigma * Sigma)); }

// ---- SRGBUtilities.fxh:27-33 ----
inline float3 LinearToSRGB(float3 RGB)
{
    float3 bGreater = float3(step(0.0031308f, RGB.x), step(0.0031308f, RGB.y), step(0.0031308f, RGB.z));
    float3 hi       = (pow3(RGB, 1.0f / 2.4f) * 1.055f) - float3(0.055f, 0.055f, 0.055f);
    return lerp(RGB * 12.92f, hi, bGreater);
}
// :5-11
inline float3 SRGBToLinear(float3 sRGB)
{
    float3 b  = float3(step(0.04045f, sRGB.x), step(0.04045f, sRGB.y), step(0.04045f, sRGB.z));
    float3 t  = (sRGB + float3(0.055f, 0.055f, 0.055f)) / 1.055f;
    float3 hi = pow3(float3(saturate(t.x), saturate(t.y), saturate(t.z)), 2.4f);
    return lerp(sRGB / 12.92f, hi, b);
}

// ---- PBR_Common.fxh (GGX subset) ----
// :107-124
inline float SmithGGXVisibilityCorrelated(float NdotL, float NdotV, float AlphaRoughness)
{
    float a2   = AlphaRoughness * AlphaRoughness;
    float GGXV = NdotL * std::sqrt(hmax(NdotV * NdotV * (1.0f - a2) + a2, 1e-7f));
    float GGXL = NdotV * std::sqrt(hmax(NdotL * NdotL * (1.0f - a2) + a2, 1e-7f));
    return 0.5f / (GGXV + GGXL);
}
// :149-176
inline float SmithGGXMasking(float NdotV, float AlphaRoughness)
{
    float a2    = AlphaRoughness * AlphaRoughness;
    float Denom = NdotV + std::sqrt(a2 + (1.0f - a2) * NdotV * NdotV);
    return 2.0f * hmax(NdotV, 0.0f) / hmax(Denom, 1e-6f);
}
// :181-195
inline float NormalDistribution_GGX(float NdotH, float AlphaRoughness)
{
    AlphaRoughness = hmax(AlphaRoughness, 1e-3f);
    float a2  = AlphaRoughness * AlphaRoughness;
    float nh2 = NdotH * NdotH;
    float f   = nh2 * a2 + (1.0f - nh2);
    return a2 / hmax(PBR_PI * f * f, 1e-9f);
}
// :278-296
inline float3 SmithGGXSampleVisibleNormalSC(float3 View, float ax, float ay, float u1, float u2)
{
    float3 V        = normalize(View * float3(ax, ay, 1.0f));
    float  Phi      = 2.0f * PBR_PI * u1;
    float  Z        = (1.0f - u2) * (1.0f + V.z) - V.z;
    float  SinTheta = std::sqrt(clampf(1.0f - Z * Z, 0.0f, 1.0f));
    float3 H        = float3(SinTheta * std::cos(Phi), SinTheta * std::sin(Phi), Z) + V;
    return normalize(float3(ax * H.x, ay * H.y, H.z));
}

// view-space normal: mul(float4(N, 0), mView).xyz
inline float3 mul_dir(float3 v, const float4x4& M) { return mul(float4(v, 0.0f), M).xyz(); }

} // namespace orc
