// ORACLE — TEST INFRASTRUCTURE ONLY.
// Render-target storage formats of the reference, as value-level quantisers (SURVEY.md Appendix B.5 "faithful" storage mode).
// The product keeps every plane in fp32 (the north-star layout) and is gated against the fp32 oracle; this mode exists to report how
// far the reference's own narrow render targets move the frame, so that nobody mistakes quantisation for a kernel difference.
// Formats and where the reference uses them:
//   R8_UNORM        AO planes (ScreenSpaceAmbientOcclusion.cpp:153-190, :269-345), SSR roughness (ScreenSpaceReflection.cpp:149-158)
//   RG8_UNORM       blue noise (PostFXContext.cpp:193-205)
//   R16_FLOAT       SSAO history length (:309-323), SSR variance / hit depth (ScreenSpaceReflection.cpp:219-250, :268-281)
//   RG16_FLOAT      closest motion (PostFXContext.cpp:275-284)
//   RGBA16_FLOAT    SSR radiance / ray direction / history / output (:197-297), TAA accumulation (TemporalAntiAliasing.cpp:101-119)
//   R11G11B10_FLOAT Bloom levels and output (Bloom.cpp:101-141)
//   R16_UNORM       depth pyramids under FEATURE_FLAG_HALF_PRECISION_DEPTH (ScreenSpaceAmbientOcclusion.cpp:101-138, :205-241)
// Conversion rules: Direct3D 11 functional spec 3.2.3: UNORM = round-to-nearest of saturate(v) * (2^n - 1), NaN -> 0; float -> smaller
// float = round-to-nearest-even, overflow -> infinity for half / clamp to the largest finite value for the unsigned 11 / 10-bit floats
// (they have no sign: negative -> 0, NaN kept).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc
{
inline float q_unorm(float v, int bits)
{
    if (!(v == v)) return 0.0f;
    const float m = float((1u << bits) - 1u);
    v             = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    return std::floor(v * m + 0.5f) / m;
}

// round-to-nearest-even of an fp32 value to a float with `mant` explicit mantissa bits and a 5-bit exponent (bias 15), returned as fp32.
// `is_signed` false: the unsigned small floats of R11G11B10 (negative -> 0, values above the largest finite one clamp to it).
inline float q_small_float(float v, int mant, bool is_signed)
{
    if (!(v == v)) return v;
    if (!is_signed && v <= 0.0f) return 0.0f;
    const float a = std::fabs(v);
    if (std::isinf(a)) return is_signed ? v : std::ldexp(float((1 << (mant + 1)) - 1), 15 - mant);
    const float max_finite = std::ldexp(float((1 << (mant + 1)) - 1), 15 - mant); // (2 - 2^-mant) * 2^15
    int         e;
    (void)std::frexp(a, &e); // a = f * 2^e, f in [0.5, 1)
    int exp2 = e - 1;        // a in [2^exp2, 2^(exp2+1))
    if (exp2 < -14) exp2 = -14; // denormal range: fixed quantum 2^(-14 - mant)
    const float quantum = std::ldexp(1.0f, exp2 - mant);
    float       q       = std::nearbyint(a / quantum) * quantum; // default rounding mode: to nearest, ties to even
    if (q > max_finite) q = is_signed ? INFINITY : max_finite;
    if (is_signed && q == max_finite + quantum) q = INFINITY;
    return v < 0.0f ? -q : q;
}
inline float q_half(float v) { return q_small_float(v, 10, true); }
inline float q_float11(float v) { return q_small_float(v, 6, false); }
inline float q_float10(float v) { return q_small_float(v, 5, false); }
} // namespace orc
