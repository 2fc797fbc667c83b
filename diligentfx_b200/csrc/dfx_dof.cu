// dfx_dof.cu — DepthOfField (sits between TAA and Bloom in the reference chain, HnPostProcessTask.cpp:899-909).
// Reference host code: PostProcess/DepthOfField/src/DepthOfField.cpp:292-331 (order), :820-1116 (bindings);
// shaders: Shaders/PostProcess/DepthOfField/private/DOF_*.fx (cited per kernel).
// Planes are fp32 like everywhere else in this library (the reference keeps the CoC planes in R16_FLOAT / R16_UNORM and the
// colour planes in RGBA16_FLOAT / R11G11B10_FLOAT); the half-size planes are (W/2) x (H/2), the dilation chain W>>k x H>>k.
#include "dfx_common.cuh"

#include <cmath>
#include <vector>

namespace dfx
{

constexpr int kMaxKernelPoints = 72; // 1 + 7 * (5 * 4 / 2) = 71 for the largest Octaweb kernel the UI allows (rings <= 5, density <= 7)
struct KernelPoints
{
    float2 p[kMaxKernelPoints];
    int    n;
};
struct Gauss13
{
    float w[13];
};

// DepthOfField.cpp:49-73 — rings from the outside in; ring i has max(density * i, 1) points, phase 0.1 * i
static KernelPoints make_kernel_points(int ring_count, int ring_density)
{
    KernelPoints k{};
    const float  radius_inc = 1.0f / (static_cast<float>(ring_count) - 1.0f);
    for (int i = ring_count - 1; i >= 0; --i)
    {
        const int   count     = std::max(ring_density * i, 1);
        const float radius    = static_cast<float>(i) * radius_inc;
        const float theta_inc = 2.0f * 3.14159265358979323846f / static_cast<float>(count);
        const float offset    = 0.1f * static_cast<float>(i);
        for (int j = 0; j < count && k.n < kMaxKernelPoints; ++j)
        {
            const float theta = offset + static_cast<float>(j) * theta_inc;
            k.p[k.n++]        = make_float2(radius * std::cos(theta), radius * std::sin(theta));
        }
    }
    return k;
}
// DepthOfField.cpp:75-91 with DOF_GAUSS_KERNEL_RADIUS 6, DOF_GAUSS_KERNEL_SIGMA 5
static Gauss13 make_gauss()
{
    Gauss13 g{};
    float   sum = 0.0f;
    for (int i = -6; i <= 6; ++i) sum += (g.w[i + 6] = std::exp(-static_cast<float>(i * i) / (2.0f * 5.0f * 5.0f)));
    for (float& v : g.w) v /= sum;
    return g;
}


// Bilinear SampleLevel (clamp) with every multiply and add rounded separately. The bokeh gathers compare an interpolated alpha
// with the centre's CoC (`SampledColor.a >= CoCFar`): a fused multiply-add that moves the interpolant by one ulp flips the
// comparison where the CoC is locally constant, and with it 1/71 of the pixel. Separate roundings make the sampler a pure
// function of (plane, uv) that the oracle (built with -ffp-contract=off) reproduces bit for bit.
template <class T> struct DofLerp;
template <> struct DofLerp<float>
{
    static __device__ __forceinline__ float run(float a, float b, float c, float d, float w00, float w10, float w01, float w11)
    {
        return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, w00), __fmul_rn(b, w10)), __fmul_rn(c, w01)), __fmul_rn(d, w11));
    }
};
template <> struct DofLerp<float4>
{
    static __device__ __forceinline__ float4 run(float4 a, float4 b, float4 c, float4 d, float w00, float w10, float w01, float w11)
    {
        return make_float4(DofLerp<float>::run(a.x, b.x, c.x, d.x, w00, w10, w01, w11), DofLerp<float>::run(a.y, b.y, c.y, d.y, w00, w10, w01, w11),
                           DofLerp<float>::run(a.z, b.z, c.z, d.z, w00, w10, w01, w11), DofLerp<float>::run(a.w, b.w, c.w, d.w, w00, w10, w01, w11));
    }
};
template <class T> __device__ __forceinline__ T dof_sample(const View<const T>& t, float u, float v)
{
    float px = __fsub_rn(__fmul_rn(u, float(t.w)), 0.5f), py = __fsub_rn(__fmul_rn(v, float(t.h)), 0.5f);
    px = __fmul_rn(floorf(__fadd_rn(__fmul_rn(px, 256.0f), 0.5f)), 1.0f / 256.0f); // 8-bit sub-texel snap
    py = __fmul_rn(floorf(__fadd_rn(__fmul_rn(py, 256.0f), 0.5f)), 1.0f / 256.0f);
    const float fx0 = floorf(px), fy0 = floorf(py);
    const int   x0 = (int)fx0, y0 = (int)fy0;
    const float fx = __fsub_rn(px, fx0), fy = __fsub_rn(py, fy0), gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
    return DofLerp<T>::run(loadc(t, x0, y0), loadc(t, x0 + 1, y0), loadc(t, x0, y0 + 1), loadc(t, x0 + 1, y0 + 1), __fmul_rn(gx, gy), __fmul_rn(fx, gy),
                           __fmul_rn(gx, fy), __fmul_rn(fx, fy));
}
// uv + offset with the offset's products rounded separately (the compiler would otherwise fold the last product into the add)
__device__ __forceinline__ float dof_offset(float k, float coc, float scale, float max_coc) { return __fmul_rn(__fmul_rn(__fmul_rn(k, scale), coc), max_coc); }

// ---------------------------------------------------------------------------------------------------------------------
// D1  DOF_ComputeCircleOfConfusion.fx:24-39 — signed CoC in [-1, 1]: < 0 near field, 0 in focus, > 0 far field
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dof_coc_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_dof_attribs A, View<const float> depth, View<float> coc,
                                                      int y0, int y1)
{
    __shared__ float lens[4]; // focus distance, f-stop, focal length, sensor width
    __shared__ CamS  cam;
    if (threadIdx.x == 0 && threadIdx.y == 0)
    {
        load_cam(cam, &cams[0]);
        lens[0] = cams[0].fFocusDistance, lens[1] = cams[0].fFStop, lens[2] = cams[0].fFocalLength, lens[3] = cams[0].fSensorWidth;
    }
    __syncthreads();
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= coc.w || y >= y1) return;
    const float d  = __ldg(&depth.at(x, y));
    const float z  = (cam.m32 - d * cam.m33) / (d * cam.m23 - cam.m22); // DepthToCameraZ, IEEE division: the CoC feeds comparisons
    const float f  = lens[2] / 1000.0f;
    const float K  = f * f / (lens[1] * (lens[0] - f));
    const float c  = K * (z - lens[0]) / fmaxf(z, 1e-4f);
    st_cs(&coc.at(x, y), fminf(fmaxf(1000.0f * c / (lens[3] * A.MaxCircleOfConfusion), -1.0f), 1.0f));
}

// ---------------------------------------------------------------------------------------------------------------------
// D2  DOF_ComputeTemporalCircleOfConfusion.fx:54-92 — history clamped to mean +- 2.5 sigma of the 3x3 neighbourhood
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dof_temporal_coc_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_dof_attribs A, View<const float> curr,
                                                               View<const float> prev, View<const float2> closest, View<float> out, int y0, int y1)
{
    __shared__ CamS cam;
    if (threadIdx.x == 0 && threadIdx.y == 0) load_cam(cam, &cams[0]);
    __syncthreads();
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float  posx = float(x) + 0.5f, posy = float(y) + 0.5f;
    const float2 mv   = __ldg(&closest.at(x, y));
    const float  ppx = posx - (mv.x * 0.5f) * cam.vw, ppy = posy - (mv.y * -0.5f) * cam.vh;
    const float  c0  = __ldg(&curr.at(x, y));
    if (!(ppx >= 0.0f && ppy >= 0.0f && ppx < cam.vw && ppy < cam.vh)) // IsInsideScreen (PostFX_Common.fxh:121-127)
    {
        st_cs(&out.at(x, y), c0);
        return;
    }
    const float cp = dof_sample(prev, __fmul_rn(ppx, cam.ivw), __fmul_rn(ppy, cam.ivh));
    float       m1 = 0.0f, m2 = 0.0f;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
        {
            const float c = sample_point_clamp(curr, (posx + float(dx)) * cam.ivw, (posy + float(dy)) * cam.ivh);
            m1 += c, m2 += c * c;
        }
    const float mean = m1 / 9.0f, var = (m2 / 9.0f) - (mean * mean), sd = sqrtf(fmaxf(var, 0.0f));
    const float lo = mean - 2.5f * sd, hi = mean + 2.5f * sd;
    st_cs(&out.at(x, y), lerpf(c0, fminf(fmaxf(cp, lo), hi), A.TemporalStabilityFactor));
}

// D3  DOF_ComputeSeparatedCircleOfConfusion.fx:5-11
__global__ void __launch_bounds__(256) dof_separated_kernel(View<const float> coc, View<float> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    if (pix.x >= out.w || pix.y >= y1) return;
    const float c = __ldg(&coc.at(pix.x, pix.y));
    st_cs(&out.at(pix.x, pix.y), c < 0.0f ? fabsf(c) : 0.0f);
}

// D4  DOF_ComputeDilationCircleOfConfusion.fx:16-52
__global__ void __launch_bounds__(256) dof_dilation_kernel(View<const float> last, View<float> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const bool wodd = last.w & 1, hodd = last.h & 1;
    const auto S    = [&](int ox, int oy) { return loadc(last, 2 * x + ox, 2 * y + oy); };
    float      m    = fmaxf(fmaxf(S(0, 0), S(0, 1)), fmaxf(S(1, 0), S(1, 1)));
    if (wodd) m = fmaxf(m, fmaxf(S(2, 0), S(2, 1)));
    if (hodd) m = fmaxf(m, fmaxf(S(0, 2), S(1, 2)));
    if (wodd && hodd) m = fmaxf(m, S(2, 2));
    out.at(x, y) = m;
}

// D5, D6  DOF_ComputeBlurredCircleOfConfusion.fx:8-28
template <bool VERTICAL>
__global__ void __launch_bounds__(256) dof_blur_kernel(View<const float> coc, View<float> out, Gauss13 g, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    float s = 0.0f;
#pragma unroll
    for (int i = -6; i <= 6; ++i) s += (VERTICAL ? loadc(coc, x, y + i) : loadc(coc, x + i, y)) * g.w[i + 6];
    out.at(x, y) = s;
}

// D7  DOF_ComputePrefilteredTexture.fx:23-52
__global__ void __launch_bounds__(256) dof_prefilter_kernel(View<const float4> color, View<const float> coc, View<const float> dilation, View<float4> out_fg,
                                                            View<float4> out_bg, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_fg.w || y >= y1) return;
    float  cmax = -kFltMax, wsum = 0.0f;
    float3 sum  = make_float3(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const int    lx = 2 * x + (i & 1), ly = 2 * y + (i >> 1);
        const float3 c = xyz(load0(color, lx, ly));
        const float  w = 1.0f / (1.0f + luminance(c)); // ComputeSDRWeight (DOF_Common.fx:14-17)
        cmax           = fmaxf(cmax, load0(coc, lx, ly));
        sum            = sum + c * w, wsum += w;
    }
    const float  fa   = dof_sample(dilation, (float(x) + 0.5f) / float(out_fg.w), (float(y) + 0.5f) / float(out_fg.h));
    const float  ba   = cmax > 0.0f ? fabsf(cmax) : 0.0f;
    const float3 mean = sum / fmaxf(wsum, 1.e-5f);
    st_cs(&out_fg.at(x, y), f4(mean, fa));
    st_cs(&out_bg.at(x, y), f4(mean, ba));
}

// D8  DOF_ComputeBokehFirstPass.fx:47-104
template <bool KARIS>
__global__ void __launch_bounds__(256) dof_bokeh_first_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_dof_attribs A, View<const float4> fg,
                                                              View<const float4> bg, View<const float4> radiance, View<float4> out_fg, View<float4> out_bg,
                                                              const __grid_constant__ KernelPoints K, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_fg.w || y >= y1) return;
    const float aspect = cams[0].f4ViewportSize[0] * cams[0].f4ViewportSize[3];
    const float u = (float(x) + 0.5f) / float(out_fg.w), v = (float(y) + 0.5f) / float(out_fg.h);
    const float cn = dof_sample(fg, u, v).w, cf = dof_sample(bg, u, v).w;
    float3      fs = make_float3(0.f, 0.f, 0.f), bs = make_float3(0.f, 0.f, 0.f);
    float       fw = 0.0f, bw = 0.0f;
    if (cn > 0.0f)
        for (int i = 0; i < K.n; ++i)
        {
            const float  sx = dof_offset(K.p[i].x, cn, 0.5f, A.MaxCircleOfConfusion), sy = dof_offset(K.p[i].y, cn, 0.5f, A.MaxCircleOfConfusion);
            const float  su = __fadd_rn(u, sx), sv = __fadd_rn(v, __fmul_rn(aspect, sy));
            const float4 s  = dof_sample(fg, su, sv);
            const float  w  = KARIS ? 1.0f + luminance(xyz(dof_sample(radiance, su, sv))) : 1.0f; // ComputeHDRWeight
            fs = fs + xyz(s) * w, fw += w;
        }
    if (cf > 0.0f)
        for (int i = 0; i < K.n; ++i)
        {
            const float  sx = dof_offset(K.p[i].x, cf, 0.5f, A.MaxCircleOfConfusion), sy = dof_offset(K.p[i].y, cf, 0.5f, A.MaxCircleOfConfusion);
            const float  su = __fadd_rn(u, sx), sv = __fadd_rn(v, __fmul_rn(aspect, sy));
            const float4 s  = dof_sample(bg, su, sv);
            float        w  = KARIS ? 1.0f + luminance(xyz(dof_sample(radiance, su, sv))) : 1.0f;
            w               = s.w >= cf ? w : 0.0f;
            bs = bs + xyz(s) * w, bw += w;
        }
    st_cs(&out_fg.at(x, y), f4(fs * (1.0f / (fw + (fw == 0.0f ? 1.0f : 0.0f))), cn));
    st_cs(&out_bg.at(x, y), f4(bs * (1.0f / (bw + (bw == 0.0f ? 1.0f : 0.0f))), cf));
}

// D9  DOF_ComputeBokehSecondPass.fx:37-85
__global__ void __launch_bounds__(256) dof_bokeh_second_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_dof_attribs A, View<const float4> fg,
                                                               View<const float4> bg, View<float4> out_fg, View<float4> out_bg,
                                                               const __grid_constant__ KernelPoints K, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_fg.w || y >= y1) return;
    const float  aspect = cams[0].f4ViewportSize[0] * cams[0].f4ViewportSize[3];
    const float  u = (float(x) + 0.5f) / float(out_fg.w), v = (float(y) + 0.5f) / float(out_fg.h);
    const float4 F0 = dof_sample(fg, u, v), B0 = dof_sample(bg, u, v);
    const float  cn = F0.w, cf = B0.w;
    float3       F = xyz(F0), B = xyz(B0);
    if (cn > 0.0f)
        for (int i = 0; i < K.n; ++i)
        {
            const float  sx = dof_offset(K.p[i].x, cn, 0.25f, A.MaxCircleOfConfusion), sy = dof_offset(K.p[i].y, cn, 0.25f, A.MaxCircleOfConfusion);
            const float4 s  = dof_sample(fg, __fadd_rn(u, sx), __fadd_rn(v, __fmul_rn(aspect, sy)));
            F               = make_float3(fmaxf(s.x, F.x), fmaxf(s.y, F.y), fmaxf(s.z, F.z));
        }
    if (cf > 0.0f)
        for (int i = 0; i < K.n; ++i)
        {
            const float  sx = dof_offset(K.p[i].x, cf, 0.25f, A.MaxCircleOfConfusion), sy = dof_offset(K.p[i].y, cf, 0.25f, A.MaxCircleOfConfusion);
            const float4 s  = dof_sample(bg, __fadd_rn(u, sx), __fadd_rn(v, __fmul_rn(aspect, sy)));
            const float  k  = s.w >= cf ? 1.0f : 0.0f;
            B               = make_float3(fmaxf(s.x * k, B.x), fmaxf(s.y * k, B.y), fmaxf(s.z * k, B.z));
        }
    st_cs(&out_fg.at(x, y), f4(F, cn));
    st_cs(&out_bg.at(x, y), f4(B, cf));
}

// D10  DOF_ComputePostfilteredTexture.fx:26-48
__global__ void __launch_bounds__(256) dof_postfilter_kernel(View<const float4> fg, View<const float4> bg, View<float4> out_fg, View<float4> out_bg, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_fg.w || y >= y1) return;
    const float u = (float(x) + 0.5f) / float(out_fg.w), v = (float(y) + 0.5f) / float(out_fg.h);
    const float tx = 1.0f / float(fg.w), ty = 1.0f / float(fg.h);
    const auto  tent = [&](const View<const float4>& t) {
        const float ul = __fadd_rn(u, __fmul_rn(tx, -0.5f)), ur = __fadd_rn(u, __fmul_rn(tx, 0.5f));
        const float vt = __fadd_rn(v, __fmul_rn(ty, -0.5f)), vb = __fadd_rn(v, __fmul_rn(ty, 0.5f));
        return (dof_sample(t, ul, vt) + dof_sample(t, ul, vb) + dof_sample(t, ur, vt) + dof_sample(t, ur, vb)) * 0.25f;
    };
    st_cs(&out_fg.at(x, y), tent(fg));
    st_cs(&out_bg.at(x, y), tent(bg));
}

// D11  DOF_ComputeCombinedTexture.fx:34-46
__global__ void __launch_bounds__(256) dof_combine_kernel(dfx_dof_attribs A, View<const float4> color, View<const float4> dnear, View<const float4> dfar,
                                                          View<float4> out, int y0, int y1)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float  u = (float(x) + 0.5f) / float(out.w), v = (float(y) + 0.5f) / float(out.h);
    const float4 src = __ldg(&color.at(x, y));
    const float4 n = dof_sample(dnear, u, v), f = dof_sample(dfar, u, v);
    const float3 s = xyz(src);
    float3       r = s + (xyz(f) - s) * smoothstepf(0.1f, 1.0f, f.w);
    r              = r + (xyz(n) - r) * smoothstepf(0.1f, 1.0f, n.w);
    st_cs(&out.at(x, y), f4(s + (r - s) * A.AlphaInterpolation, src.w));
}

} // namespace dfx

using namespace dfx;

#define DFX_GRID(w, rows) dim3 block(32, 8), grid(div_up(w, 32), div_up(rows.y1 - rows.y0, 8))
#define DFX_ROWS_OF(view, rows)                                  \
    DFX_REQUIRE(rows_ok(rows, (view).h), "bad row range (rows of the output plane)"); \
    if (rows.y1 == rows.y0) return DFX_OK

extern "C" dfx_status dfx_pass_dof_coc(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_dof_attribs* attribs, const dfx_plane* depth,
                                       const dfx_plane* out_coc, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_coc");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    DFX_VIEW(float, o, out_coc, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(d, o);
    DFX_ROWS_OF(o, rows);
    DFX_GRID(o.w, rows);
    dof_coc_kernel<<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, d, o, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_coc_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_temporal_coc(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_dof_attribs* attribs, const dfx_plane* curr_coc,
                                                const dfx_plane* prev_coc, const dfx_plane* closest_motion, const dfx_plane* out_coc, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_temporal_coc");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const float, c, curr_coc, DFX_FORMAT_R32F);
    DFX_VIEW(const float, p, prev_coc, DFX_FORMAT_R32F);
    DFX_VIEW(const float2, m, closest_motion, DFX_FORMAT_RG32F);
    DFX_VIEW(float, o, out_coc, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(c, p);
    DFX_SAME_SIZE(c, m);
    DFX_SAME_SIZE(c, o);
    DFX_ROWS_OF(o, rows);
    DFX_GRID(o.w, rows);
    dof_temporal_coc_kernel<<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, c, p, m, o, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_temporal_coc_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_separated_coc(void* stream, const dfx_plane* coc, const dfx_plane* out, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_separated_coc");
    DFX_VIEW(const float, c, coc, DFX_FORMAT_R32F);
    DFX_VIEW(float, o, out, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(c, o);
    DFX_ROWS_OF(o, rows);
    DFX_GRID(o.w, rows);
    dof_separated_kernel<<<grid, block, 0, as_stream(stream)>>>(c, o, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_separated_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_dilation(void* stream, const dfx_plane* last, const dfx_plane* out, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_dilation");
    DFX_VIEW(const float, l, last, DFX_FORMAT_R32F);
    DFX_VIEW(float, o, out, DFX_FORMAT_R32F);
    DFX_REQUIRE(o.w == (l.w >> 1) && o.h == (l.h >> 1), "a dilation level is (width >> 1) x (height >> 1) of the previous one");
    DFX_ROWS_OF(o, rows);
    DFX_GRID(o.w, rows);
    dof_dilation_kernel<<<grid, block, 0, as_stream(stream)>>>(l, o, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_dilation_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_blur_coc(void* stream, const dfx_plane* coc, int32_t vertical, const dfx_plane* out, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_blur_coc");
    DFX_VIEW(const float, c, coc, DFX_FORMAT_R32F);
    DFX_VIEW(float, o, out, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(c, o);
    DFX_REQUIRE(c.p != o.p, "the blur is not in-place: ping-pong through the intermediate plane");
    DFX_ROWS_OF(o, rows);
    DFX_GRID(o.w, rows);
    const Gauss13 g = make_gauss();
    if (vertical)
        dof_blur_kernel<true><<<grid, block, 0, as_stream(stream)>>>(c, o, g, rows.y0, rows.y1);
    else
        dof_blur_kernel<false><<<grid, block, 0, as_stream(stream)>>>(c, o, g, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_blur_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_prefilter(void* stream, const dfx_plane* color, const dfx_plane* coc, const dfx_plane* dilation, const dfx_plane* out_fg,
                                             const dfx_plane* out_bg, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_prefilter");
    DFX_VIEW(const float4, c, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float, k, coc, DFX_FORMAT_R32F);
    DFX_VIEW(const float, dl, dilation, DFX_FORMAT_R32F);
    DFX_VIEW(float4, of, out_fg, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, ob, out_bg, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(c, k);
    DFX_SAME_SIZE(of, ob);
    DFX_REQUIRE(of.w == c.w / 2 && of.h == c.h / 2, "the prefiltered planes are width/2 x height/2 of the colour plane");
    DFX_ROWS_OF(of, rows);
    DFX_GRID(of.w, rows);
    dof_prefilter_kernel<<<grid, block, 0, as_stream(stream)>>>(c, k, dl, of, ob, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_prefilter_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_bokeh(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_dof_attribs* attribs, uint32_t flags, int32_t second_pass,
                                         const dfx_plane* fg, const dfx_plane* bg, const dfx_plane* radiance, const dfx_plane* out_fg, const dfx_plane* out_bg,
                                         dfx_rows rows)
{
    DFX_PROFILE(stream, second_pass ? "dof_bokeh_second" : "dof_bokeh_first");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const float4, f, fg, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, b, bg, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, of, out_fg, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, ob, out_bg, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(f, b);
    DFX_SAME_SIZE(f, of);
    DFX_SAME_SIZE(f, ob);
    DFX_REQUIRE(f.p != of.p && b.p != ob.p, "the gather is not in-place");
    DFX_ROWS_OF(of, rows);
    DFX_GRID(of.w, rows);
    if (second_pass)
    {
        const KernelPoints K = make_kernel_points(3, 5); // DOF_BOKEH_KERNEL_SMALL_RING_COUNT / _DENSITY
        dof_bokeh_second_kernel<<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, f, b, of, ob, K, rows.y0, rows.y1);
    }
    else
    {
        DFX_REQUIRE(attribs->BokehKernelRingCount >= 2 && attribs->BokehKernelRingCount <= 5 && attribs->BokehKernelRingDensity >= 2 && attribs->BokehKernelRingDensity <= 7,
                    "Octaweb kernel: ring count 2..5, ring density 2..7");
        const KernelPoints K = make_kernel_points(attribs->BokehKernelRingCount, attribs->BokehKernelRingDensity);
        if (flags & DFX_DOF_FEATURE_FLAG_ENABLE_KARIS_INVERSE)
        {
            DFX_VIEW(const float4, r, radiance, DFX_FORMAT_RGBA32F);
            dof_bokeh_first_kernel<true><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, f, b, r, of, ob, K, rows.y0, rows.y1);
        }
        else
        {
            View<const float4> r{nullptr, 0, 0, 0};
            dof_bokeh_first_kernel<false><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, f, b, r, of, ob, K, rows.y0, rows.y1);
        }
    }
    DFX_LAUNCHED("dof_bokeh_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_postfilter(void* stream, const dfx_plane* fg, const dfx_plane* bg, const dfx_plane* out_fg, const dfx_plane* out_bg, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_postfilter");
    DFX_VIEW(const float4, f, fg, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, b, bg, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, of, out_fg, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, ob, out_bg, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(f, b);
    DFX_SAME_SIZE(f, of);
    DFX_SAME_SIZE(f, ob);
    DFX_REQUIRE(f.p != of.p && b.p != ob.p, "the filter is not in-place");
    DFX_ROWS_OF(of, rows);
    DFX_GRID(of.w, rows);
    dof_postfilter_kernel<<<grid, block, 0, as_stream(stream)>>>(f, b, of, ob, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_postfilter_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_dof_combine(void* stream, const dfx_dof_attribs* attribs, const dfx_plane* color, const dfx_plane* dof_near, const dfx_plane* dof_far,
                                           const dfx_plane* out, dfx_rows rows)
{
    DFX_PROFILE(stream, "dof_combine");
    DFX_REQUIRE(attribs, "null argument");
    DFX_VIEW(const float4, c, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, n, dof_near, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, f, dof_far, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, o, out, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(c, o);
    DFX_SAME_SIZE(n, f);
    DFX_ROWS_OF(o, rows);
    DFX_GRID(o.w, rows);
    dof_combine_kernel<<<grid, block, 0, as_stream(stream)>>>(*attribs, c, n, f, o, rows.y0, rows.y1);
    DFX_LAUNCHED("dof_combine_kernel");
    return DFX_OK;
}
