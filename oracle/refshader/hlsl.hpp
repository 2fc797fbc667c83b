// REFERENCE-SHADER RUNNER — TEST INFRASTRUCTURE ONLY (never on the product path; see oracle/refshader/README.md).
//
// A small HLSL execution environment for g++: enough of the language's vector types, swizzles, intrinsics and texture
// objects that the reference's own pixel-shader sources (Shaders/PostProcess/**/*.fx, Shaders/Common/**/*.fxh, read in
// place from /root/reference by build_ref.py) compile as C++ and run one invocation per pixel on the CPU. Everything
// here is this repository's code; nothing is taken from the reference. What it models:
//   * float / int / uint / bool vectors of 2-4 components with xyzw / rgba swizzles (readable and writable),
//     float3x3 / float4x4 with the row-vector mul() convention the reference uses (mul(v, M) == v * M),
//   * the intrinsics the PostProcess shaders call (see the grep in README.md), all in fp32,
//   * Texture2D<T>::Load / SampleLevel / GetDimensions and SamplerState with the Direct3D rules the oracle also states in
//     oracle_tex.h: Load out of bounds returns 0, missing channels read (0,0,0,1), bilinear positions snapped to 1/256 texel,
//     clamp or border(0) addressing, nearest-mip selection for point-mip samplers.
// The fixed-function parts (sampler arithmetic, render-target formats = fp32 planes) are therefore the same MODEL as the
// oracle's; what this runner adds is that the per-pixel shader math is the reference's own source text.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <memory>
#include <ucontext.h>
#include <vector>

namespace hlsl
{
typedef unsigned int uint;

template <class T, int N> struct vec;

// A swizzle is a view of some components of the vector whose storage it shares (it lives in the vector's union).
template <class T, int M, int... I> struct swz
{
    T d[M];
    typedef vec<T, int(sizeof...(I))> V;
    operator V() const { return V(d[I]...); }
    swz& operator=(const V& v)
    {
        int k = 0;
        ((d[I] = v.d[k++]), ...);
        return *this;
    }
    swz& operator=(const swz& o) { return *this = V(o); }
    template <int M2, int... J> swz& operator=(const swz<T, M2, J...>& o) { return *this = V(o); }
    swz& operator+=(const V& v) { return *this = V(*this) + v; }
    swz& operator-=(const V& v) { return *this = V(*this) - v; }
    swz& operator*=(const V& v) { return *this = V(*this) * v; }
    swz& operator/=(const V& v) { return *this = V(*this) / v; }
};

#define HLSL_VEC_COMMON(N)                                                                              \
    vec()                                                                                               \
    {                                                                                                   \
        for (int i = 0; i < N; ++i) d[i] = T();                                                         \
    }                                                                                                   \
    vec(const vec& o)                                                                                   \
    {                                                                                                   \
        for (int i = 0; i < N; ++i) d[i] = o.d[i];                                                      \
    }                                                                                                   \
    vec& operator=(const vec& o)                                                                        \
    {                                                                                                   \
        for (int i = 0; i < N; ++i) d[i] = o.d[i];                                                      \
        return *this;                                                                                   \
    }                                                                                                   \
    vec(T s)                                                                                            \
    {                                                                                                   \
        for (int i = 0; i < N; ++i) d[i] = s;                                                           \
    }                                                                                                   \
    template <class U> explicit vec(const vec<U, N>& o)                                                 \
    {                                                                                                   \
        for (int i = 0; i < N; ++i) d[i] = T(o.d[i]);                                                   \
    }                                                                                                   \
    template <class U, int M, int... I> explicit vec(const swz<U, M, I...>& s) : vec(s.operator vec<U, N>()) {} \
    T&       operator[](int i) { return d[i]; }                                                         \
    const T& operator[](int i) const { return d[i]; }

template <class T> struct vec<T, 2>
{
    union
    {
        T d[2];
        struct
        {
            T x, y;
        };
        struct
        {
            T r, g;
        };
#include "swizzles_2.inc"
    };
    HLSL_VEC_COMMON(2)
    vec(T a, T b) : d{a, b} {}
};
template <class T> struct vec<T, 3>
{
    union
    {
        T d[3];
        struct
        {
            T x, y, z;
        };
        struct
        {
            T r, g, b;
        };
#include "swizzles_3.inc"
    };
    HLSL_VEC_COMMON(3)
    vec(T a, T b, T c) : d{a, b, c} {}
    vec(const vec<T, 2>& a, T c) : d{a.x, a.y, c} {}
    vec(T a, const vec<T, 2>& b) : d{a, b.x, b.y} {}
    // HLSL converts element types inside constructors (int3(float2, 0) truncates)
    template <class U> vec(const vec<U, 2>& a, T c) : d{T(a.x), T(a.y), c} {}
    template <class U, int M, int I0, int I1> vec(const swz<U, M, I0, I1>& a, T c) : d{T(a.d[I0]), T(a.d[I1]), c} {}
};
template <class T> struct vec<T, 4>
{
    union
    {
        T d[4];
        struct
        {
            T x, y, z, w;
        };
        struct
        {
            T r, g, b, a;
        };
#include "swizzles_4.inc"
    };
    HLSL_VEC_COMMON(4)
    vec(T a_, T b_, T c, T e) : d{a_, b_, c, e} {}
    vec(const vec<T, 2>& p, T c, T e) : d{p.x, p.y, c, e} {}
    vec(T a_, T b_, const vec<T, 2>& p) : d{a_, b_, p.x, p.y} {}
    vec(const vec<T, 2>& p, const vec<T, 2>& q) : d{p.x, p.y, q.x, q.y} {}
    vec(const vec<T, 3>& p, T e) : d{p.x, p.y, p.z, e} {}
    vec(T a_, const vec<T, 3>& p) : d{a_, p.x, p.y, p.z} {}
    template <class U> vec(const vec<U, 3>& p, T e) : d{T(p.x), T(p.y), T(p.z), e} {}
    template <class U, int M, int I0, int I1, int I2> vec(const swz<U, M, I0, I1, I2>& p, T e) : d{T(p.d[I0]), T(p.d[I1]), T(p.d[I2]), e} {}
    template <class U> vec(const vec<U, 2>& p, T c, T e) : d{T(p.x), T(p.y), c, e} {}
    template <class U, int M, int I0, int I1> vec(const swz<U, M, I0, I1>& p, T c, T e) : d{T(p.d[I0]), T(p.d[I1]), c, e} {}
};

typedef vec<float, 2> float2;
typedef vec<float, 3> float3;
typedef vec<float, 4> float4;
typedef vec<int, 2>   int2;
typedef vec<int, 3>   int3;
typedef vec<int, 4>   int4;
typedef vec<uint, 2>  uint2;
typedef vec<uint, 3>  uint3;
typedef vec<uint, 4>  uint4;
typedef vec<bool, 2>  bool2;
typedef vec<bool, 3>  bool3;
typedef vec<bool, 4>  bool4;

// 4-byte boolean of the constant-buffer structures (BOOL in ShaderDefinitions.fxh is `bool` on the HLSL side, which
// occupies a 32-bit register there): keeps the structures' layout identical to the C structs of include/dfx_b200.h.
struct bool32
{
    int v;
    bool32(bool b = false) : v(b ? 1 : 0) {}
    operator bool() const { return v != 0; }
};

// ---- operators: non-template overloads per concrete type, so that swizzles and scalars convert implicitly ----
#define HLSL_BINOP(V, N, OP)                                          \
    inline V operator OP(const V& a, const V& b)                      \
    {                                                                 \
        V r;                                                          \
        for (int i = 0; i < N; ++i) r.d[i] = a.d[i] OP b.d[i];        \
        return r;                                                     \
    }                                                                 \
    inline V& operator OP##=(V& a, const V& b) { return a = a OP b; }
#define HLSL_CMPOP(V, N, OP)                                          \
    inline vec<bool, N> operator OP(const V& a, const V& b)           \
    {                                                                 \
        vec<bool, N> r;                                               \
        for (int i = 0; i < N; ++i) r.d[i] = a.d[i] OP b.d[i];        \
        return r;                                                     \
    }
#define HLSL_ARITH(V, N)                                              \
    HLSL_BINOP(V, N, +) HLSL_BINOP(V, N, -) HLSL_BINOP(V, N, *) HLSL_BINOP(V, N, /) \
    HLSL_CMPOP(V, N, <) HLSL_CMPOP(V, N, >) HLSL_CMPOP(V, N, <=) HLSL_CMPOP(V, N, >=) HLSL_CMPOP(V, N, ==) HLSL_CMPOP(V, N, !=) \
    inline V operator-(const V& a)                                    \
    {                                                                 \
        V r;                                                          \
        for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];                 \
        return r;                                                     \
    }
#define HLSL_BITS(V, N) HLSL_BINOP(V, N, &) HLSL_BINOP(V, N, |) HLSL_BINOP(V, N, ^) HLSL_BINOP(V, N, <<) HLSL_BINOP(V, N, >>) HLSL_BINOP(V, N, %)

HLSL_ARITH(float2, 2) HLSL_ARITH(float3, 3) HLSL_ARITH(float4, 4)
HLSL_ARITH(int2, 2) HLSL_ARITH(int3, 3) HLSL_ARITH(int4, 4)
HLSL_ARITH(uint2, 2) HLSL_ARITH(uint3, 3) HLSL_ARITH(uint4, 4)
HLSL_BITS(int2, 2) HLSL_BITS(int3, 3) HLSL_BITS(int4, 4)
HLSL_BITS(uint2, 2) HLSL_BITS(uint3, 3) HLSL_BITS(uint4, 4)

#define HLSL_BOOLV(N)                                                                 \
    inline bool any(const vec<bool, N>& a)                                            \
    {                                                                                 \
        bool r = false;                                                               \
        for (int i = 0; i < N; ++i) r = r || a.d[i];                                  \
        return r;                                                                     \
    }                                                                                 \
    inline bool all(const vec<bool, N>& a)                                            \
    {                                                                                 \
        bool r = true;                                                                \
        for (int i = 0; i < N; ++i) r = r && a.d[i];                                  \
        return r;                                                                     \
    }
HLSL_BOOLV(2) HLSL_BOOLV(3) HLSL_BOOLV(4)

// ---- scalar intrinsics (fp32; the C library's float functions) ----
inline float abs(float x) { return std::fabs(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float rsqrt(float x) { return 1.0f / std::sqrt(x); }
inline float rcp(float x) { return 1.0f / x; }
inline float floor(float x) { return std::floor(x); }
inline float ceil(float x) { return std::ceil(x); }
inline float round(float x) { return std::nearbyint(x); } // HLSL round(): to nearest even
inline float trunc(float x) { return std::trunc(x); }
inline float frac(float x) { return x - std::floor(x); }
inline float exp(float x) { return std::exp(x); }
inline float exp2(float x) { return std::exp2(x); }
inline float log(float x) { return std::log(x); }
inline float log2(float x) { return std::log2(x); }
inline float log10(float x) { return std::log10(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float tan(float x) { return std::tan(x); }
inline float asin(float x) { return std::asin(x); }
inline float acos(float x) { return std::acos(x); }
inline float atan(float x) { return std::atan(x); }
inline float atan2(float y, float x) { return std::atan2(y, x); }
inline float pow(float x, float y) { return std::pow(x, y); }
inline float fmod(float x, float y) { return std::fmod(x, y); }
inline float saturate(float x) { return std::fmin(std::fmax(x, 0.0f), 1.0f); } // NaN -> 0 (Direct3D: min / max return the non-NaN operand)
inline float sign(float x) { return float((x > 0.0f) - (x < 0.0f)); }
inline float step(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
inline float min(float a, float b) { return std::fmin(a, b); } // D3D min/max return the non-NaN operand
inline float max(float a, float b) { return std::fmax(a, b); }
inline int   min(int a, int b) { return a < b ? a : b; }
inline int   max(int a, int b) { return a > b ? a : b; }
inline uint  min(uint a, uint b) { return a < b ? a : b; }
inline uint  max(uint a, uint b) { return a > b ? a : b; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int   clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline uint  clamp(uint x, uint lo, uint hi) { return min(max(x, lo), hi); }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float mad(float a, float b, float c) { return a * b + c; }
inline float smoothstep(float a, float b, float x)
{
    float t = saturate((x - a) / (b - a));
    return t * t * (3.0f - 2.0f * t);
}
inline void  sincos(float x, float& s, float& c) { s = std::sin(x), c = std::cos(x); }
// Screen-space derivatives: the four pixels of a 2x2 quad run in lockstep (quad_run below); ddx / ddy exchange the operand
// between them and return the fine difference inside the quad (right - left, bottom - top). Outside quad_run they return 0.
float ddx(float v);
float ddy(float v);
inline bool  isnan(float x) { return std::isnan(x); }
inline bool  isinf(float x) { return std::isinf(x); }
inline uint  asuint(float x)
{
    uint u;
    std::memcpy(&u, &x, 4);
    return u;
}
inline uint  asuint(uint x) { return x; }
inline float asfloat(uint u)
{
    float x;
    std::memcpy(&x, &u, 4);
    return x;
}
inline float asfloat(float x) { return x; }
inline uint  reversebits(uint v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}
inline uint countbits(uint v)
{
    uint n = 0;
    for (; v; v &= v - 1) ++n;
    return n;
}

// ---- component-wise vector forms ----
#define HLSL_MAP1(V, N, F)                                     \
    inline V F(const V& a)                                     \
    {                                                          \
        V r;                                                   \
        for (int i = 0; i < N; ++i) r.d[i] = F(a.d[i]);        \
        return r;                                              \
    }
#define HLSL_MAP2(V, N, F)                                     \
    inline V F(const V& a, const V& b)                         \
    {                                                          \
        V r;                                                   \
        for (int i = 0; i < N; ++i) r.d[i] = F(a.d[i], b.d[i]); \
        return r;                                              \
    }
#define HLSL_MAP3(V, N, F)                                             \
    inline V F(const V& a, const V& b, const V& c)                     \
    {                                                                  \
        V r;                                                           \
        for (int i = 0; i < N; ++i) r.d[i] = F(a.d[i], b.d[i], c.d[i]); \
        return r;                                                      \
    }
#define HLSL_FLOATV(V, N)                                                                                               \
    HLSL_MAP1(V, N, abs) HLSL_MAP1(V, N, sqrt) HLSL_MAP1(V, N, rsqrt) HLSL_MAP1(V, N, rcp) HLSL_MAP1(V, N, floor)        \
    HLSL_MAP1(V, N, ceil) HLSL_MAP1(V, N, round) HLSL_MAP1(V, N, trunc) HLSL_MAP1(V, N, frac) HLSL_MAP1(V, N, exp)       \
    HLSL_MAP1(V, N, exp2) HLSL_MAP1(V, N, log) HLSL_MAP1(V, N, log2) HLSL_MAP1(V, N, log10) HLSL_MAP1(V, N, sin)         \
    HLSL_MAP1(V, N, cos) HLSL_MAP1(V, N, saturate) HLSL_MAP1(V, N, sign) HLSL_MAP1(V, N, ddx) HLSL_MAP1(V, N, ddy)       \
    HLSL_MAP2(V, N, pow) HLSL_MAP2(V, N, min) HLSL_MAP2(V, N, max) HLSL_MAP2(V, N, step) HLSL_MAP2(V, N, fmod)           \
    HLSL_MAP2(V, N, atan2) HLSL_MAP3(V, N, clamp) HLSL_MAP3(V, N, lerp) HLSL_MAP3(V, N, smoothstep) HLSL_MAP3(V, N, mad) \
    inline float dot(const V& a, const V& b)                                                                            \
    {                                                                                                                   \
        float s = a.d[0] * b.d[0];                                                                                      \
        for (int i = 1; i < N; ++i) s += a.d[i] * b.d[i];                                                               \
        return s;                                                                                                       \
    }                                                                                                                   \
    inline float length(const V& a) { return std::sqrt(dot(a, a)); }                                                    \
    inline float distance(const V& a, const V& b) { return length(a - b); }                                             \
    inline V     normalize(const V& a) { return a / length(a); } /* same statement of the intrinsic as oracle_math.h */                                               \
    inline V     reflect(const V& i, const V& n) { return i - n * (2.0f * dot(i, n)); }                                 \
    inline V     lerp(const V& a, const V& b, float t) { return a + (b - a) * t; }                                                 \
    inline V     lerp(const V& a, const V& b, const vec<bool, N>& t) { return lerp(a, b, V(t)); } /* bool -> 0.0 / 1.0 */
HLSL_FLOATV(float2, 2) HLSL_FLOATV(float3, 3) HLSL_FLOATV(float4, 4)
#define HLSL_INTV(V, N) HLSL_MAP2(V, N, min) HLSL_MAP2(V, N, max) HLSL_MAP3(V, N, clamp)
HLSL_INTV(int2, 2) HLSL_INTV(int3, 3) HLSL_INTV(int4, 4) HLSL_INTV(uint2, 2) HLSL_INTV(uint3, 3) HLSL_INTV(uint4, 4)

inline float3 cross(const float3& a, const float3& b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// ---- matrices: rows; mul(v, M) is the row vector times M (the reference writes mul(float4(Pos, 1.0), g_Camera.mViewProj)) ----
struct float4x4
{
    float4        r[4];
    float4&       operator[](int i) { return r[i]; }
    const float4& operator[](int i) const { return r[i]; }
};
struct float3x3
{
    float3 r[3];
    float3x3() {}
    float3x3(const float3& a, const float3& b, const float3& c) : r{a, b, c} {}
    explicit float3x3(const float4x4& m) : r{float3(m.r[0].x, m.r[0].y, m.r[0].z), float3(m.r[1].x, m.r[1].y, m.r[1].z), float3(m.r[2].x, m.r[2].y, m.r[2].z)} {}
    float3&       operator[](int i) { return r[i]; }
    const float3& operator[](int i) const { return r[i]; }
};
inline float4 mul(const float4& v, const float4x4& m) { return m.r[0] * v.x + m.r[1] * v.y + m.r[2] * v.z + m.r[3] * v.w; }
inline float4 mul(const float4x4& m, const float4& v) { return float4(dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v), dot(m.r[3], v)); }
inline float3 mul(const float3& v, const float3x3& m) { return m.r[0] * v.x + m.r[1] * v.y + m.r[2] * v.z; }
inline float3 mul(const float3x3& m, const float3& v) { return float3(dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)); }
inline float4x4 mul(const float4x4& a, const float4x4& b)
{
    float4x4 c;
    for (int i = 0; i < 4; ++i) c.r[i] = mul(a.r[i], b);
    return c;
}
inline float3x3 transpose(const float3x3& m) { return float3x3(float3(m.r[0].x, m.r[1].x, m.r[2].x), float3(m.r[0].y, m.r[1].y, m.r[2].y), float3(m.r[0].z, m.r[1].z, m.r[2].z)); }

// ---- textures ----
struct SamplerState
{
    int linear     = 0; // min/mag filter: 0 point, 1 linear
    int mip_linear = 0; // mip filter: 0 nearest, 1 linear
    int border     = 0; // address mode: 0 clamp, 1 border (colour 0)
};

template <class T> struct texel_traits
{
    typedef T scalar;
    enum
    {
        N = 1
    };
    static T make(const scalar* c) { return c[0]; }
};
template <class S, int M> struct texel_traits<vec<S, M>>
{
    typedef S scalar;
    enum
    {
        N = M
    };
    static vec<S, M> make(const scalar* c)
    {
        vec<S, M> v;
        for (int i = 0; i < M; ++i) v.d[i] = c[i];
        return v;
    }
};

constexpr int kMaxMips = 16;

template <class T = float4> struct Texture2D
{
    typedef typename texel_traits<T>::scalar S;
    const S* mip[kMaxMips] = {};
    int      w[kMaxMips] = {}, h[kMaxMips] = {};
    int      levels = 0, ch = 0;

    void bind(int level, const void* data, int width, int height, int channels)
    {
        mip[level] = static_cast<const S*>(data), w[level] = width, h[level] = height, ch = channels;
        levels = std::max(levels, level + 1);
    }
    // in-bounds texel; channels the resource does not have read (0, 0, 0, 1)
    T fetch(int m, int x, int y) const
    {
        S        c[4] = {S(0), S(0), S(0), S(1)};
        const S* p = mip[m] + (size_t(y) * size_t(w[m]) + size_t(x)) * size_t(ch);
        for (int i = 0; i < ch && i < 4; ++i) c[i] = p[i];
        return texel_traits<T>::make(c);
    }
    T load(int m, int x, int y) const
    {
        if (m < 0 || m >= levels || x < 0 || y < 0 || x >= w[m] || y >= h[m]) return T(S(0));
        return fetch(m, x, y);
    }
    T Load(const int3& p) const { return load(p.z, p.x, p.y); }
    T Load(const uint3& p) const { return load(int(p.z), int(p.x), int(p.y)); }
    T clamped(int m, int x, int y) const { return fetch(m, std::min(std::max(x, 0), w[m] - 1), std::min(std::max(y, 0), h[m] - 1)); }

    T sample_mip(const SamplerState& s, const float2& uv, int m) const
    {
        if (!s.linear)
        {
            const int x = int(std::floor(uv.x * float(w[m]))), y = int(std::floor(uv.y * float(h[m])));
            return s.border ? load(m, x, y) : clamped(m, x, y);
        }
        float px = uv.x * float(w[m]) - 0.5f, py = uv.y * float(h[m]) - 0.5f;
        px = std::floor(px * 256.0f + 0.5f) * (1.0f / 256.0f); // 8 fractional bits of the fixed-function sampler
        py = std::floor(py * 256.0f + 0.5f) * (1.0f / 256.0f);
        const float fx0 = std::floor(px), fy0 = std::floor(py);
        const int   x0 = int(fx0), y0 = int(fy0);
        const float fx = px - fx0, fy = py - fy0;
        const T     t00 = s.border ? load(m, x0, y0) : clamped(m, x0, y0), t10 = s.border ? load(m, x0 + 1, y0) : clamped(m, x0 + 1, y0);
        const T     t01 = s.border ? load(m, x0, y0 + 1) : clamped(m, x0, y0 + 1), t11 = s.border ? load(m, x0 + 1, y0 + 1) : clamped(m, x0 + 1, y0 + 1);
        return t00 * ((1.0f - fx) * (1.0f - fy)) + t10 * (fx * (1.0f - fy)) + t01 * ((1.0f - fx) * fy) + t11 * (fx * fy);
    }
    T SampleLevel(const SamplerState& s, const float2& uv, float lod) const
    {
        if (!s.mip_linear || levels == 1) return sample_mip(s, uv, std::min(std::max(int(std::floor(lod + 0.5f)), 0), levels - 1));
        const float l  = std::min(std::max(lod, 0.0f), float(levels - 1));
        const int   m0 = int(std::floor(l)), m1 = std::min(m0 + 1, levels - 1);
        const float f  = l - float(m0);
        return sample_mip(s, uv, m0) * (1.0f - f) + sample_mip(s, uv, m1) * f;
    }
    // Sample(): implicit-derivative LOD. Only single-mip look-up tables are sampled this way on this path -> mip 0.
    T Sample(const SamplerState& s, const float2& uv) const { return sample_mip(s, uv, 0); }
    template <class U> void GetDimensions(U& W, U& H) const { W = U(w[0]), H = U(h[0]); }
    template <class U, class L> void GetDimensions(uint m, U& W, U& H, L& n) const { W = U(w[m]), H = U(h[m]), n = L(levels); }
};

// Cube maps are declared by PBR_Shading.fxh (environment lighting) but never sampled on the PostProcess path: the type
// exists so that those helper functions compile; sampling one returns 0.
template <class T = float4> struct TextureCube
{
    T SampleLevel(const SamplerState&, const float3&, float) const { return T(0.0f); }
    T Sample(const SamplerState&, const float3&) const { return T(0.0f); }
};

// ---- the full-screen pass: one pixel-shader invocation per target pixel, rows split over host threads ----
// A persistent pool (one per library: an inline function's static object is shared by all translation units of the .so), rows
// handed out in small chunks, so that a pass costs no thread creation - with 128 host cores and ~80 passes per frame, creating
// the threads per pass would cost more than shading the pixels.
struct row_pool
{
    struct worker
    {
        std::condition_variable cv;
        unsigned long long      wanted = 0, seen = 0; // generation this worker is asked to join / has joined (guarded by m)
    };
    std::mutex                                m;
    std::condition_variable                   cv_done;
    std::vector<std::unique_ptr<worker>>      workers;
    const std::function<void(int, int)>*      fn = nullptr;
    std::atomic<int>                          next{0};
    int                                       height = 0, chunk = 1, pending = 0;
    unsigned long long                        generation = 0;

    explicit row_pool(int n)
    {
        for (int i = 0; i < n; ++i)
        {
            workers.emplace_back(new worker);
            worker* w = workers.back().get();
            std::thread([this, w] {
                std::unique_lock<std::mutex> lk(m);
                for (;;)
                {
                    w->cv.wait(lk, [&] { return w->wanted != w->seen; });
                    w->seen = w->wanted;
                    lk.unlock();
                    drain();
                    lk.lock();
                    if (--pending == 0) cv_done.notify_one();
                }
            }).detach();
        }
    }
    void drain()
    {
        for (int y = next.fetch_add(chunk); y < height; y = next.fetch_add(chunk)) (*fn)(y, std::min(height, y + chunk));
    }
    // `threads` row workers in total: the caller is one of them, and only the helpers this pass uses are woken (waking all 128 workers
    // of a big host for every one of the ~80 passes of a frame costs more than a small pass's pixels)
    void run(int h, int threads, const std::function<void(int, int)>& f)
    {
        const int helpers = std::min(threads - 1, int(workers.size()));
        {
            std::lock_guard<std::mutex> lk(m);
            fn = &f, height = h, chunk = std::max(1, h / (8 * threads)), pending = helpers;
            next = 0, ++generation;
            for (int i = 0; i < helpers; ++i) workers[i]->wanted = generation;
        }
        for (int i = 0; i < helpers; ++i) workers[i]->cv.notify_one();
        drain();
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};
inline row_pool& the_row_pool()
{
    static row_pool* p = new row_pool(int(std::max(1u, std::thread::hardware_concurrency())));
    return *p;
}
inline void for_rows(int height, int threads, const std::function<void(int, int)>& fn)
{
    static std::mutex           one_pass_at_a_time;
    row_pool&                   pool = the_row_pool();
    threads = std::min(std::min(threads, int(pool.workers.size())), height / 2);
    if (threads <= 1 || height < 8) return fn(0, height);
    std::lock_guard<std::mutex> lk(one_pass_at_a_time);
    pool.run(height, threads, fn);
}

// ---- 2x2 quads in lockstep, for ddx / ddy: each of the four lanes is a coroutine on its own stack. Lane l shades pixel
// (2*qx + (l & 1), 2*qy + (l >> 1)). A lane that calls ddx / ddy publishes its operand and yields; once every running lane
// has yielded the values are snapshotted and the lanes resume (Direct3D requires derivatives in quad-uniform control flow).
struct quad_state
{
    ucontext_t                      sched, lane[4];
    std::vector<char>               stack[4];
    float                           val[4] = {}, snap[4] = {};
    bool                            finished[4] = {};
    int                             current = -1;
    const std::function<void(int)>* body = nullptr;
};
inline quad_state*& quad_tls()
{
    static thread_local quad_state* q = nullptr;
    return q;
}
inline void quad_lane_entry()
{
    quad_state* q = quad_tls();
    const int   l = q->current;
    (*q->body)(l);
    q->finished[l] = true; // returning activates uc_link == the scheduler
}
inline float quad_exchange(float v, int axis)
{
    quad_state* q = quad_tls();
    if (!q || q->current < 0) return 0.0f;
    const int l = q->current;
    q->val[l]   = v;
    swapcontext(&q->lane[l], &q->sched);
    q = quad_tls();
    return axis == 0 ? q->snap[l | 1] - q->snap[l & ~1] : q->snap[l | 2] - q->snap[l & ~2];
}
inline float ddx(float v) { return quad_exchange(v, 0); }
inline float ddy(float v) { return quad_exchange(v, 1); }
inline void  quad_run(const std::function<void(int)>& body)
{
    static thread_local quad_state q;
    constexpr size_t               kStack = 512 * 1024;
    q.body = &body;
    for (int l = 0; l < 4; ++l)
    {
        if (q.stack[l].empty()) q.stack[l].resize(kStack);
        getcontext(&q.lane[l]);
        q.lane[l].uc_stack.ss_sp   = q.stack[l].data();
        q.lane[l].uc_stack.ss_size = kStack;
        q.lane[l].uc_link          = &q.sched;
        makecontext(&q.lane[l], quad_lane_entry, 0);
        q.finished[l] = false, q.val[l] = 0.0f;
    }
    quad_tls() = &q;
    for (;;)
    {
        bool running = false;
        for (int l = 0; l < 4; ++l)
            if (!q.finished[l])
            {
                q.current = l;
                swapcontext(&q.sched, &q.lane[l]); // runs lane l up to its next derivative, or to its end
                running = running || !q.finished[l];
            }
        if (!running) break;
        for (int l = 0; l < 4; ++l) q.snap[l] = q.val[l];
    }
    q.current  = -1;
    quad_tls() = nullptr;
}

} // namespace hlsl
