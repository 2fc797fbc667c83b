// cuemu — DEVELOPMENT TOOL (see cuda_runtime.h). IEEE binary16 -> binary32 for the transfer-format kernels.
#pragma once
#include "cuda_runtime.h"
struct __half
{
    unsigned short bits;
};
struct __half2
{
    __half x, y;
};
inline float __half2float(__half h)
{
    const unsigned s = (h.bits >> 15) & 1u, e = (h.bits >> 10) & 0x1Fu, m = h.bits & 0x3FFu;
    float          v;
    if (e == 0) v = std::ldexp(float(m), -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = std::ldexp(float(m | 0x400u), int(e) - 25);
    return s ? -v : v;
}
inline float2 __half22float2(__half2 h) { return make_float2(__half2float(h.x), __half2float(h.y)); }
