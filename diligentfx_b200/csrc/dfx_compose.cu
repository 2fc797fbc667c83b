// dfx_compose.cu — the full compose step of the reference integration and the table it needs.
//
//   Hydrogent/shaders/HnPostProcess.psh:145-185: with Opacity = Color.a,
//     Color.rgb += (GetSpecularIBL_GGX(SrfInfo, IBLInfo, SSR.rgb) - SpecularIBL.rgb) * SSR.a * SSRScale * Opacity
//     Color.rgb *= lerp(1, AO, SSAOScale * Opacity)
//   i.e. the screen-space reflection is re-weighted by the split-sum BRDF of the surface and EXCHANGED for the image-based
//   specular term the renderer had already added (PBR_Shading.fxh:220-302 with USE_IBL_MULTIPLE_SCATTERING = 1, :429-451).
//   The split sum reads a pre-integrated GGX table: Shaders/PBR/private/PrecomputeBRDF.psh:10-48 (PBR_Renderer.cpp:548-625
//   builds it once, 512 x 512, 512 samples). dfx_pass_compose (dfx_bloom_taa_tonemap.cu) stays as the reduced form for
//   callers without the IBL / base-colour planes.
#include "dfx_common.cuh"

namespace dfx
{

// ---------------------------------------------------------------------------------------------------------------------
// PrecomputeBRDF.psh:10-48 + PBR_PrecomputeCommon.fxh:10-39 (N = +z, so the tangent frame is the identity: UpVector = x for
// |N.z| >= 0.999, TangentX = normalize(cross(x, z)) = -y ... written out below exactly as the shader composes it)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) brdf_lut_kernel(View<float2> lut, unsigned num_samples)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= lut.w || y >= lut.h) return;
    const float NoV = (float(x) + 0.5f) / float(lut.w), rough = (float(y) + 0.5f) / float(lut.h);
    const float3 V  = make_float3(sqrtf(1.0f - NoV * NoV), 0.0f, NoV);
    const float3 N  = make_float3(0.0f, 0.0f, 1.0f);
    const float  alpha = rough * rough, a2 = alpha * alpha;
    // ImportanceSampleGGX's frame for N = (0, 0, 1): |N.z| >= 0.999 -> UpVector = (1, 0, 0)
    const float3 up = make_float3(1.0f, 0.0f, 0.0f);
    const float3 tx = normalize(cross(up, N)), ty = cross(N, tx);
    float        A = 0.0f, B = 0.0f;
    for (unsigned i = 0u; i < num_samples; ++i)
    {
        const float xi_x = float(i) / float(num_samples), xi_y = float(__brev(i)) * 2.3283064365386963e-10f; // Hammersley2D
        const float phi  = 2.0f * 3.141592653589793f * xi_x;
        const float ct   = sqrtf(saturate((1.0f - xi_y) / (1.0f + (a2 - 1.0f) * xi_y)));
        const float st   = sqrtf(saturate(1.0f - ct * ct));
        const float3 Hl  = make_float3(st * cosf(phi), st * sinf(phi), ct);
        const float3 H   = tx * Hl.x + ty * Hl.y + N * Hl.z;
        const float3 L   = 2.0f * dot(V, H) * H - V;
        const float  NoL = saturate(L.z), NoH = saturate(H.z), VoH = saturate(dot(V, H));
        if (NoL > 0.0f)
        {
            // SmithGGXVisibilityCorrelated (PBR_Common.fxh:107-124)
            const float ggxv = NoL * sqrtf(fmaxf(NoV * NoV * (1.0f - a2) + a2, 1e-7f));
            const float ggxl = NoV * sqrtf(fmaxf(NoL * NoL * (1.0f - a2) + a2, 1e-7f));
            const float vis  = 0.5f / (ggxv + ggxl);
            const float gvis = 4.0f * vis * VoH * NoL / NoH;
            const float fc   = powf(1.0f - VoH, 5.0f);
            A += (1.0f - fc) * gvis;
            B += fc * gvis;
        }
    }
    lut.at(x, y) = make_float2(A / float(num_samples), B / float(num_samples));
}

// ---------------------------------------------------------------------------------------------------------------------
// HnPostProcess.psh:145-185
// ---------------------------------------------------------------------------------------------------------------------
struct ComposeCam
{
    CamS c;
    Mat4 vp_inv;
};

__global__ void __launch_bounds__(256) compose_ibl_kernel(const dfx_camera_attribs* __restrict__ cams, View<const float4> color, View<const float4> ssr,
                                                          View<const float> ao, View<const float4> spec_ibl, View<const float4> normal,
                                                          View<const float4> base_color, View<const float4> material, View<const float2> lut,
                                                          float ssr_scale, float ssao_scale, View<float4> out, int y0, int y1)
{
    __shared__ ComposeCam S;
    if (threadIdx.x == 0 && threadIdx.y == 0) load_cam(S.c, &cams[0]), load_mat(S.vp_inv, cams[0].mViewProjInv);
    __syncthreads();
    const CamS&   cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    const float4 C = __ldg(&color.at(x, y));
    float3       c = xyz(C);
    const float  opacity = C.w;
    const float  s_ssr = ssr_scale * opacity;
    if (ssr.p && s_ssr > 0.0f)
    {
        const float4 ibl = __ldg(&spec_ibl.at(x, y)), r = __ldg(&ssr.at(x, y)), bc = __ldg(&base_color.at(x, y)), m = __ldg(&material.at(x, y));
        const float3 n   = xyz(__ldg(&normal.at(x, y)));
        const float  rough = saturate(m.x), metal = saturate(m.y);
        const float3 f0v = make_float3(0.04f, 0.04f, 0.04f);
        const float3 r0  = f0v + metal * (xyz(bc) - f0v); // lerp(f0, BaseColor, Metallic)
        const float3 wp  = inv_project_position((float(x) + 0.5f) * cam.ivw, (float(y) + 0.5f) * cam.ivh, 0.5f, S.vp_inv);
        const float3 v   = normalize(make_float3(cam.px, cam.py, cam.pz) - wp);
        const float  ndv = saturate(dot(n, v));
        const float2 pre = sample_linear_clamp(lut, ndv, rough);
        const float  omr = 1.0f - rough;
        const float3 r90 = make_float3(fmaxf(omr, r0.x), fmaxf(omr, r0.y), fmaxf(omr, r0.z));
        const float  t = fminf(fmaxf(1.0f - ndv, 0.0f), 1.0f), t2 = t * t, t5 = t2 * t2 * t;
        const float3 ks = r0 + (r90 - r0) * t5; // SchlickReflection
        const float3 spec = xyz(r) * (ks * pre.x + make_float3(pre.y, pre.y, pre.y));
        c = c + (spec - xyz(ibl)) * r.w * s_ssr; // left to right, as HnPostProcess.psh:170 multiplies
    }
    const float s_ao = ssao_scale * opacity;
    if (ao.p && s_ao > 0.0f) c = c * lerpf(1.0f, __ldg(&ao.at(x, y)), s_ao);
    st_cs(&out.at(x, y), f4(c, C.w));
}

} // namespace dfx

using namespace dfx;

extern "C" dfx_status dfx_pass_precompute_brdf_lut(void* stream, uint32_t num_samples, const dfx_plane* out_lut)
{
    DFX_PROFILE(stream, "precompute_brdf_lut");
    DFX_VIEW(float2, l, out_lut, DFX_FORMAT_RG32F);
    DFX_REQUIRE(num_samples > 0, "num_samples must be positive");
    dim3 block(32, 8), grid(div_up(l.w, 32), div_up(l.h, 8));
    brdf_lut_kernel<<<grid, block, 0, as_stream(stream)>>>(l, num_samples);
    DFX_LAUNCHED("brdf_lut_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_compose_ibl(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_plane* color, const dfx_plane* ssr,
                                           const dfx_plane* ao, const dfx_plane* specular_ibl, const dfx_plane* normal, const dfx_plane* base_color,
                                           const dfx_plane* material, const dfx_plane* brdf_lut, float ssr_scale, float ssao_scale, const dfx_plane* out,
                                           dfx_rows rows)
{
    DFX_PROFILE(stream, "compose_ibl");
    DFX_REQUIRE(cameras_dev, "null argument");
    DFX_VIEW(const float4, c, color, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, o, out, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(c, o);
    View<const float4> s{nullptr, 0, 0, 0}, ibl{nullptr, 0, 0, 0}, n{nullptr, 0, 0, 0}, bc{nullptr, 0, 0, 0}, m{nullptr, 0, 0, 0};
    View<const float>  a{nullptr, 0, 0, 0};
    View<const float2> l{nullptr, 0, 0, 0};
    if (ssr)
    {
        DFX_REQUIRE(specular_ibl && normal && base_color && material && brdf_lut, "the SSR term needs the specular IBL, normal, base colour, material and BRDF table planes");
        DFX_REQUIRE(make_view<const float4>(ssr, DFX_FORMAT_RGBA32F, s) && make_view<const float4>(specular_ibl, DFX_FORMAT_RGBA32F, ibl) &&
                        make_view<const float4>(normal, DFX_FORMAT_RGBA32F, n) && make_view<const float4>(base_color, DFX_FORMAT_RGBA32F, bc) &&
                        make_view<const float4>(material, DFX_FORMAT_RGBA32F, m) && make_view<const float2>(brdf_lut, DFX_FORMAT_RG32F, l),
                    "bad plane (null, wrong format, pitch or alignment)");
        DFX_SAME_SIZE(c, s);
        DFX_SAME_SIZE(c, ibl);
        DFX_SAME_SIZE(c, n);
        DFX_SAME_SIZE(c, bc);
        DFX_SAME_SIZE(c, m);
    }
    if (ao)
    {
        DFX_REQUIRE(make_view<const float>(ao, DFX_FORMAT_R32F, a), "bad plane 'ao'");
        DFX_SAME_SIZE(c, a);
    }
    DFX_REQUIRE(rows_ok(rows, c.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    dim3 block(32, 8), grid(div_up(c.w, 32), div_up(rows.y1 - rows.y0, 8));
    compose_ibl_kernel<<<grid, block, 0, as_stream(stream)>>>(cameras_dev, c, s, a, ibl, n, bc, m, l, ssr_scale, ssao_scale, o, rows.y0, rows.y1);
    DFX_LAUNCHED("compose_ibl_kernel");
    return DFX_OK;
}
