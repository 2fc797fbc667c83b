// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
//
// Parity status: PINNED AGAINST THE REFERENCE'S OWN SHADERS, pass by pass. The reference (DiligentFX) ships no golden
// vectors or numerical tests for its PostProcess shaders and its C++ cannot be built here (DiligentCore
// is not vendored), but its pixel shaders can be run: oracle/refshader compiles the HLSL sources where they lie under
// /root/reference for the CPU (oracle/_ref/librefshaders.so) and tests/test_reference_shaders.py holds every function of
// this oracle to the shader it restates, BIT FOR BIT, on the same inputs, in every variant; tests/golden/
// reference_shaders_49x27.npz carries those shader outputs to machines without the reference.
// What that does not cover: the host-side sequencing (which plane feeds which pass, clears, ping-pong: restated from the
// .cpp files, cited per function) and the fixed-function model both sides share (oracle_tex.h). The analytic known-answer
// tests the shaders imply stay in tests/test_oracle_kats.py.
//
// HLSL scalar/vector vocabulary restated for plain C++ (no SIMD, no FMA contraction: build with
// -ffp-contract=off). Conventions follow SURVEY.md Appendix B: row-vector mul(v, M), M[r][c],
// D3D/Vulkan NDC (y up, z in [0,1]), lerp(a,b,t) = a + t*(b-a), min/max with IEEE minNum/maxNum NaN rules.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc
{

using uint = uint32_t;

struct float2
{
    float x = 0, y = 0;
    float2() = default;
    float2(float x_, float y_) : x(x_), y(y_) {}
};
struct float3
{
    float x = 0, y = 0, z = 0;
    float3() = default;
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(float2 v, float z_) : x(v.x), y(v.y), z(z_) {}
    float2 xy() const { return {x, y}; }
};
struct float4
{
    float x = 0, y = 0, z = 0, w = 0;
    float4() = default;
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(float3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float3 xyz() const { return {x, y, z}; }
    float2 xy() const { return {x, y}; }
};
struct int2
{
    int x = 0, y = 0;
    int2() = default;
    int2(int x_, int y_) : x(x_), y(y_) {}
};

// Row-major 4x4, element (r, c) = m[r][c]; mul(v, M) treats v as a row vector (reference: DiligentCore BasicMath,
// MATRIX_ELEMENT(M, r, c) == M[r][c]).
struct float4x4
{
    float m[4][4] = {};
};

#define ORC_OP2(T, OP)                                                                              \
    inline T operator OP(T a, T b) { return T(a.x OP b.x, a.y OP b.y); }                           \
    inline T operator OP(T a, float b) { return T(a.x OP b, a.y OP b); }                           \
    inline T operator OP(float a, T b) { return T(a OP b.x, a OP b.y); }
#define ORC_OP3(T, OP)                                                                              \
    inline T operator OP(T a, T b) { return T(a.x OP b.x, a.y OP b.y, a.z OP b.z); }               \
    inline T operator OP(T a, float b) { return T(a.x OP b, a.y OP b, a.z OP b); }                 \
    inline T operator OP(float a, T b) { return T(a OP b.x, a OP b.y, a OP b.z); }
#define ORC_OP4(T, OP)                                                                              \
    inline T operator OP(T a, T b) { return T(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); }   \
    inline T operator OP(T a, float b) { return T(a.x OP b, a.y OP b, a.z OP b, a.w OP b); }       \
    inline T operator OP(float a, T b) { return T(a OP b.x, a OP b.y, a OP b.z, a OP b.w); }
ORC_OP2(float2, +) ORC_OP2(float2, -) ORC_OP2(float2, *) ORC_OP2(float2, /)
ORC_OP3(float3, +) ORC_OP3(float3, -) ORC_OP3(float3, *) ORC_OP3(float3, /)
ORC_OP4(float4, +) ORC_OP4(float4, -) ORC_OP4(float4, *) ORC_OP4(float4, /)
#undef ORC_OP2
#undef ORC_OP3
#undef ORC_OP4
inline float2 operator-(float2 a) { return {-a.x, -a.y}; }
inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
inline float2& operator+=(float2& a, float2 b) { a = a + b; return a; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
inline int2 operator+(int2 a, int2 b) { return {a.x + b.x, a.y + b.y}; }
inline int2 operator-(int2 a, int2 b) { return {a.x - b.x, a.y - b.y}; }

// ---- scalar intrinsics ----
inline float saturate(float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float frac(float v) { return v - std::floor(v); }
inline float rcp(float v) { return 1.0f / v; }
inline float sign(float v) { return float((v > 0.0f) - (v < 0.0f)); }
inline float step(float edge, float v) { return v >= edge ? 1.0f : 0.0f; }
inline float clampf(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }
inline int   clampi(int v, int lo, int hi) { return std::min(std::max(v, lo), hi); }
inline float smoothstep(float a, float b, float x)
{
    float t = saturate((x - a) / (b - a));
    return t * t * (3.0f - 2.0f * t);
}
// HLSL int(float) with D3D ftoi semantics: truncate toward zero, NaN -> 0, out-of-range saturates
// (CUDA's float->int conversion behaves the same way; plain C++ casts are UB there).
inline int ftoi(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return -2147483647 - 1;
    return int(v);
}
inline uint  asuint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline float asfloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
// HLSL min/max: if one operand is NaN the other is returned (D3D functional spec) == fminf/fmaxf.
inline float hmin(float a, float b) { return std::fmin(a, b); }
inline float hmax(float a, float b) { return std::fmax(a, b); }

// ---- vector intrinsics ----
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(float2 a) { return std::sqrt(dot(a, a)); }
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
inline float distance(float3 a, float3 b) { return length(a - b); }
inline float3 normalize(float3 a) { return a / length(a); }
inline float3 cross(float3 a, float3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float3 reflect(float3 i, float3 n) { return i - 2.0f * dot(n, i) * n; }
inline float2 lerp(float2 a, float2 b, float t) { return a + t * (b - a); }
inline float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
inline float4 lerp(float4 a, float4 b, float t) { return a + t * (b - a); }
inline float3 lerp(float3 a, float3 b, float3 t) { return a + t * (b - a); }
inline float2 saturate(float2 v) { return {saturate(v.x), saturate(v.y)}; }
inline float4 saturate(float4 v) { return {saturate(v.x), saturate(v.y), saturate(v.z), saturate(v.w)}; }
inline float2 floor2(float2 v) { return {std::floor(v.x), std::floor(v.y)}; }
inline float3 max3(float3 a, float b) { return {hmax(a.x, b), hmax(a.y, b), hmax(a.z, b)}; }
inline float4 max4(float4 a, float b) { return {hmax(a.x, b), hmax(a.y, b), hmax(a.z, b), hmax(a.w, b)}; }
inline float4 clamp4(float4 v, float4 lo, float4 hi)
{
    return {clampf(v.x, lo.x, hi.x), clampf(v.y, lo.y, hi.y), clampf(v.z, lo.z, hi.z), clampf(v.w, lo.w, hi.w)};
}
inline float3 sqrt3(float3 v) { return {std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)}; }
inline float4 sqrt4(float4 v) { return {std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z), std::sqrt(v.w)}; }
inline float3 pow3(float3 v, float e) { return {std::pow(v.x, e), std::pow(v.y, e), std::pow(v.z, e)}; }
inline float3 sign3(float3 v) { return {sign(v.x), sign(v.y), sign(v.z)}; }

// mul(float4 row-vector, M)
inline float4 mul(float4 v, const float4x4& M)
{
    return {
        v.x * M.m[0][0] + v.y * M.m[1][0] + v.z * M.m[2][0] + v.w * M.m[3][0],
        v.x * M.m[0][1] + v.y * M.m[1][1] + v.z * M.m[2][1] + v.w * M.m[3][1],
        v.x * M.m[0][2] + v.y * M.m[1][2] + v.z * M.m[2][2] + v.w * M.m[3][2],
        v.x * M.m[0][3] + v.y * M.m[1][3] + v.z * M.m[2][3] + v.w * M.m[3][3]};
}

inline float4x4 matmul(const float4x4& A, const float4x4& B)
{
    float4x4 R;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
        {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += A.m[r][k] * B.m[k][c];
            R.m[r][c] = s;
        }
    return R;
}

} // namespace orc
