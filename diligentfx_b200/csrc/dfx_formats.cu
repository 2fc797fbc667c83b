// dfx_formats.cu — ingest / egress in the reference's native render-target formats.
//
// The kernels of this library keep every plane in fp32. The reference's G-buffer is narrower
// (Hydrogent/src/Tasks/HnBeginFrameTask.cpp:63-69: scene colour RGBA16_FLOAT, motion RG16_FLOAT, normal RGBA16_FLOAT,
// material RG8_UNORM, depth D32_FLOAT) and its final target is an 8-bit sRGB swap chain. When frames cross PCIe the
// narrow formats are what travels: 30 B/px in and 4 B/px out instead of 64 and 16. Widening is exact (every half and
// every UNORM8 value is an fp32 value), so a chain fed through dfx_pass_unpack_plane computes exactly what it computes
// on the widened fp32 planes; the 8-bit pack follows the D3D UNORM rule (saturate, * 255, + 0.5, truncate; NaN -> 0).
#include <cuda_fp16.h>

#include "dfx_common.cuh"

namespace dfx
{

__global__ void __launch_bounds__(256) unpack_rgba16f_kernel(View<const uint2> src, View<float4> dst, int y0, int y1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dst.w || y >= y1) return;
    const uint2   v  = __ldg(&src.at(x, y));
    const __half2 lo = *reinterpret_cast<const __half2*>(&v.x), hi = *reinterpret_cast<const __half2*>(&v.y);
    const float2  a = __half22float2(lo), b = __half22float2(hi);
    dst.at(x, y)    = make_float4(a.x, a.y, b.x, b.y);
}

__global__ void __launch_bounds__(256) unpack_rg16f_kernel(View<const uint32_t> src, View<float2> dst, int y0, int y1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dst.w || y >= y1) return;
    const uint32_t v = __ldg(&src.at(x, y));
    dst.at(x, y)     = __half22float2(*reinterpret_cast<const __half2*>(&v));
}

__global__ void __launch_bounds__(256) unpack_rg8_kernel(View<const uchar2> src, View<float4> dst, int y0, int y1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dst.w || y >= y1) return;
    const uchar2 v = __ldg(&src.at(x, y));
    dst.at(x, y)   = make_float4(float(v.x) / 255.0f, float(v.y) / 255.0f, 0.0f, 0.0f); // UNORM -> float: c / 255, correctly rounded
}

__device__ __forceinline__ unsigned unorm8(float v) { return (unsigned)__fmaf_rn(fminf(fmaxf(v, 0.0f), 1.0f), 255.0f, 0.5f); }

__global__ void __launch_bounds__(256) pack_rgba8_kernel(View<const float4> src, View<uchar4> dst, int y0, int y1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dst.w || y >= y1) return;
    const float4 v = __ldg(&src.at(x, y));
    dst.at(x, y)   = make_uchar4((unsigned char)unorm8(v.x), (unsigned char)unorm8(v.y), (unsigned char)unorm8(v.z), (unsigned char)unorm8(v.w));
}

} // namespace dfx

using namespace dfx;

#define DFX_GRID(w, rows) dim3 block(32, 8), grid(div_up(w, 32), div_up(rows.y1 - rows.y0, 8))

extern "C" dfx_status dfx_pass_unpack_plane(void* stream, const dfx_plane* src, const dfx_plane* dst, dfx_rows rows)
{
    DFX_PROFILE(stream, "unpack_gbuffer");
    DFX_REQUIRE(src && dst, "null argument");
    DFX_REQUIRE(src->width == dst->width && src->height == dst->height, "plane size mismatch: src vs dst");
    DFX_REQUIRE(rows_ok(rows, dst->height), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(dst->width, rows);
    switch (src->format)
    {
        case DFX_FORMAT_RGBA16F:
        {
            DFX_VIEW(const uint2, s, src, DFX_FORMAT_RGBA16F);
            DFX_VIEW(float4, d, dst, DFX_FORMAT_RGBA32F);
            unpack_rgba16f_kernel<<<grid, block, 0, as_stream(stream)>>>(s, d, rows.y0, rows.y1);
            break;
        }
        case DFX_FORMAT_RG16F:
        {
            DFX_VIEW(const uint32_t, s, src, DFX_FORMAT_RG16F);
            DFX_VIEW(float2, d, dst, DFX_FORMAT_RG32F);
            unpack_rg16f_kernel<<<grid, block, 0, as_stream(stream)>>>(s, d, rows.y0, rows.y1);
            break;
        }
        case DFX_FORMAT_RG8U:
        {
            DFX_VIEW(const uchar2, s, src, DFX_FORMAT_RG8U);
            DFX_VIEW(float4, d, dst, DFX_FORMAT_RGBA32F);
            unpack_rg8_kernel<<<grid, block, 0, as_stream(stream)>>>(s, d, rows.y0, rows.y1);
            break;
        }
        default: return set_error(DFX_ERR_INVALID_ARG, "unpack: source format %d is not RGBA16F / RG16F / RG8U", src->format);
    }
    DFX_LAUNCHED("unpack_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_pack_ldr8(void* stream, const dfx_plane* src, const dfx_plane* dst, dfx_rows rows)
{
    DFX_PROFILE(stream, "pack_ldr8");
    DFX_VIEW(const float4, s, src, DFX_FORMAT_RGBA32F);
    DFX_VIEW(uchar4, d, dst, DFX_FORMAT_RGBA8U);
    DFX_SAME_SIZE(s, d);
    DFX_REQUIRE(rows_ok(rows, d.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(d.w, rows);
    pack_rgba8_kernel<<<grid, block, 0, as_stream(stream)>>>(s, d, rows.y0, rows.y1);
    DFX_LAUNCHED("pack_rgba8_kernel");
    return DFX_OK;
}
