// dfx_ssr.cu — ScreenSpaceReflection passes S1, S2, S4-S7 as sm_100a kernels.
// Reference host code: PostProcess/ScreenSpaceReflection/src/ScreenSpaceReflection.cpp:777-1104;
// shaders: Shaders/PostProcess/ScreenSpaceReflection/private/SSR_*.fx (cited per kernel).
// The D16 stencil mask of the reference is an R8U plane here: 1 = reflection sample, 0 = masked out. Consumers that the
// reference draws depth-tested against the mask simply skip masked pixels (their targets keep their previous content).
#include "dfx_common.cuh"
#include "dfx_pyramid.cuh"
#ifndef DFX_INTERSECT_V2
#    define DFX_INTERSECT_V2 0
#endif

namespace dfx
{

DFX_HD bool is_reflection_sample(float rough, float depth, float thr, int rev) { return rough <= thr && !is_background(depth, rev); } // SSR_Common.fxh:57-60

__host__ __device__ inline int ssr_mip_row(int y, int m, int full_h, int mip_h) { return y >= full_h ? mip_h : (y >> m); }

// S1 (Hi-Z pyramid, SSR_ComputeHierarchicalDepthBuffer.fx:30-73): HizOp in dfx_pyramid.cuh — tile / tail / per-level kernels.

// ---------------------------------------------------------------------------------------------------------------------
// S2: reflection mask + roughness extraction — SSR_ComputeStencilMaskAndExtractRoughness.fx:13-40
// ---------------------------------------------------------------------------------------------------------------------
template <int MFMT>
__global__ void __launch_bounds__(256) ssr_mask_kernel(dfx_ssr_attribs A, Tex4T<MFMT> material, View<const float> depth,
                                                       View<float> roughness, View<uint8_t> mask, int y0, int y1, int rev)
{
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= depth.w || y >= y1) return;
    const float4 m = material.ld(x, y);
    float r = A.RoughnessChannel == 0u ? m.x : A.RoughnessChannel == 1u ? m.y : A.RoughnessChannel == 2u ? m.z : A.RoughnessChannel == 3u ? m.w : 0.0f;
    if (!A.IsRoughnessPerceptual) r = sqrtf(r);
    const bool pass = is_reflection_sample(r, __ldg(&depth.at(x, y)), A.RoughnessThreshold, rev);
    mask.at(x, y) = pass ? 1 : 0;
    if (pass) roughness.at(x, y) = r; // masked-out texels keep their stale value, like the reference's un-cleared target
}

// ---------------------------------------------------------------------------------------------------------------------
// S3 (half resolution only): downsampled mask — closest depth and largest roughness of the 2x2 (+ odd row / column)
// footprint pass IsReflectionSample — SSR_ComputeDownsampledStencilMask.fx:13-61
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ssr_downsample_mask_kernel(dfx_ssr_attribs A, View<const float> roughness, View<const float> depth,
                                                                  View<uint8_t> mask, int y0, int y1, int rev)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= mask.w || y >= y1) return;
    const bool wodd = depth.w & 1, hodd = depth.h & 1;
    float      md = rev ? 0.0f : 1.0f, mr = 0.0f;
    const auto upd = [&](int ox, int oy) {
        const float d = loadc(depth, 2 * x + ox, 2 * y + oy);
        md            = rev ? fmaxf(md, d) : fminf(md, d);
        mr            = fmaxf(mr, loadc(roughness, 2 * x + ox, 2 * y + oy));
    };
    upd(0, 0), upd(1, 0), upd(0, 1), upd(1, 1);
    if (wodd) upd(2, 0), upd(2, 1);
    if (hodd) upd(0, 2), upd(1, 2);
    if (wodd && hodd) upd(2, 2);
    mask.at(x, y) = is_reflection_sample(mr, md, A.RoughnessThreshold, rev) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// S4: stochastic GGX ray + Hi-Z march + hit validation — SSR_ComputeIntersection.fx:281-325
// ---------------------------------------------------------------------------------------------------------------------
struct IntersectCam
{
    CamS c;
    Mat4 view, view_inv, proj;
};

struct HizView
{
    View<const float> lv[DFX_MAX_MIPS];
    int               levels;
};

__device__ __forceinline__ float hiz_load(const HizView& h, int x, int y, int mip)
{
    if (mip < 0 || mip >= h.levels) return 0.0f;
    return load0(h.lv[mip], x, y);
}
struct __align__(16) HizLevel
{
    const float* p;
    int          pitch;
    int          wh; // width | height << 16  (0 for a level the pyramid does not have: every Load returns 0)
    // The march's per-level screen resolution and its reciprocal (SSR_ComputeIntersection.fx:150-152, 179-183 keep them as
    // running products by 2 / 0.5; powers of two scale exactly, so sw * 2^-k and 1 / (sw * 2^-k) are the same floats).
    float resx, resy, irx, iry;
};
static_assert(sizeof(HizLevel) == 32, "two 128-bit shared loads per level");
__device__ __forceinline__ float hiz_load(const HizLevel& L, int x, int y)
{
    const unsigned w = (unsigned)L.wh & 0xFFFFu, hgt = (unsigned)L.wh >> 16;
    // planes are < 2^31 texels: 32-bit index arithmetic (one IMAD + one IMAD.WIDE instead of the 64-bit expansion)
    return ((unsigned)x < w && (unsigned)y < hgt) ? __ldg(L.p + (unsigned)(y * L.pitch + x)) : 0.0f;
}
__device__ __forceinline__ float hiz_load(const HizLevel* lvl, int x, int y, int mip) { return hiz_load(lvl[mip & (DFX_MAX_MIPS - 1)], x, y); }

// PBR_Common.fxh:181-195
DFX_HD float ndf_ggx(float NdotH, float a)
{
    a         = fmaxf(a, 1e-3f);
    float a2  = a * a;
    float nh2 = NdotH * NdotH;
    float f   = nh2 * a2 + (1.0f - nh2);
    return a2 / fmaxf(3.141592653589793f * f * f, 1e-9f);
}
// PBR_Common.fxh:149-176
DFX_HD float smith_ggx_masking(float NdotV, float a)
{
    float a2 = a * a;
    float dn = NdotV + sqrtf(a2 + (1.0f - a2) * NdotV * NdotV);
    return 2.0f * fmaxf(NdotV, 0.0f) / fmaxf(dn, 1e-6f);
}
// PBR_Common.fxh:107-124
DFX_HD float smith_ggx_visibility_correlated(float NdotL, float NdotV, float a)
{
    float a2   = a * a;
    float ggxv = NdotL * sqrtf(fmaxf(NdotV * NdotV * (1.0f - a2) + a2, 1e-7f));
    float ggxl = NdotV * sqrtf(fmaxf(NdotL * NdotL * (1.0f - a2) + a2, 1e-7f));
    return 0.5f / (ggxv + ggxl);
}

DFX_HD float edge_vignette(float hx, float hy, float sw, float sh) // :192-197
{
    float fx = 0.05f * (sh / sw), fy = 0.05f;
    float bx = smoothstepf(0.0f, fx, hx) * (1.0f - smoothstepf(1.0f - fx, 1.0f, hx));
    float by = smoothstepf(0.0f, fy, hy) * (1.0f - smoothstepf(1.0f - fy, 1.0f, hy));
    return bx * by;
}

// Frame split into row strips over several GPUs (dfx_pass_ssr_intersect_peer): base pointers of the planes the march reads
// beyond its own strip, per owning rank. Rank r owns the full-res rows [row_begin[r], row_begin[r+1]); boundaries are
// multiples of 64 so row y of level k is owned by the owner of full-res row y << k.
struct PeerArgs
{
    const float*  hiz[DFX_MAX_MIPS][DFX_MAX_PEERS];
    const float4* color[DFX_MAX_PEERS];
    const float4* normal[DFX_MAX_PEERS];
    int           row_begin[DFX_MAX_PEERS + 1];
    int           count;
};
struct PeerTables
{
    const float*  hiz[DFX_MAX_MIPS * DFX_MAX_PEERS];
    const float4* color[DFX_MAX_PEERS];
    const float4* normal[DFX_MAX_PEERS];
    uint8_t       owner[kPeerMaxBlocks];
};
struct NoPeerTables
{
};
__device__ __forceinline__ float hiz_load(const HizLevel& L, const PeerTables& P, int x, int y, int mip)
{
    const unsigned w = (unsigned)L.wh & 0xFFFFu, hgt = (unsigned)L.wh >> 16;
    if (!((unsigned)x < w && (unsigned)y < hgt)) return 0.0f;
    const float* p = P.hiz[(mip & (DFX_MAX_MIPS - 1)) * DFX_MAX_PEERS + P.owner[(y << mip) >> kPeerBlockShift]];
    return __ldg(p + (unsigned)(y * L.pitch + x));
}
__device__ __forceinline__ float hiz_load(const HizLevel& L, const NoPeerTables&, int x, int y, int) { return hiz_load(L, x, y); }
// Load of a full-res RGBA plane at the hit texel (0 out of bounds): from the owner of row y.
template <bool NORMAL, class V> __device__ __forceinline__ float4 hit_load0(const V& v, const PeerTables& P, int x, int y) // peer planes are RGBA32F (host-checked)
{
    if (!((unsigned)x < (unsigned)v.w && (unsigned)y < (unsigned)v.h)) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* base = (NORMAL ? P.normal : P.color)[P.owner[y >> kPeerBlockShift]];
    return __ldg(base + (unsigned)(y * (v.pitch >> 4) + x));
}
template <bool NORMAL, class V> __device__ __forceinline__ float4 hit_load0(const V& v, const NoPeerTables&, int x, int y) { return load0(v, x, y); }

template <bool PREV_FRAME, bool PEER, bool REV, bool G16>
__global__ void __launch_bounds__(256, DFX_OCC_INTERSECT) ssr_intersect_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssr_attribs A,
                                                            TexRGBA<G16> color, TexRGBA<G16> normal, View<const float> roughness,
                                                            View<const uint8_t> mask, View<const float2> noise, HizView hiz,
                                                            TexRG<G16> motion, View<float4> out_rad, View<float4> out_dir, int y0, int y1, int half,
                                                            const __grid_constant__ typename std::conditional<PEER, PeerArgs, NoPeerTables>::type peer_args)
{
    __shared__ IntersectCam S;
    // Hi-Z level table in shared memory: the march picks a level per iteration, and one 16-byte LDS (pointer, pitch, packed
    // size) is cheaper than four register-indexed constant-bank loads of the kernel-parameter struct.
    __shared__ HizLevel lvl[DFX_MAX_MIPS];
    __shared__ typename std::conditional<PEER, PeerTables, NoPeerTables>::type PT;
    if (threadIdx.x == 0 && threadIdx.y == 0)
    {
        load_cam(S.c, &cams[0]);
        load_mat(S.view, cams[0].mView);
        load_mat(S.view_inv, cams[0].mViewInv);
        load_mat(S.proj, cams[0].mProj);
    }
    if (threadIdx.y == 1 && threadIdx.x < DFX_MAX_MIPS)
    {
        const int i = threadIdx.x;
        const float rx = cams[0].f4ViewportSize[0] * (1.0f / float(1 << i)), ry = cams[0].f4ViewportSize[1] * (1.0f / float(1 << i));
        HizLevel    L  = i < hiz.levels ? HizLevel{hiz.lv[i].p, hiz.lv[i].pitch, hiz.lv[i].w | (hiz.lv[i].h << 16)} : HizLevel{nullptr, 0, 0};
        L.resx = rx, L.resy = ry, L.irx = 1.0f / rx, L.iry = 1.0f / ry;
        lvl[i] = L;
    }
    if constexpr (PEER)
    {
        const int tid = threadIdx.y * blockDim.x + threadIdx.x; // 256 threads
        if (tid < DFX_MAX_MIPS * DFX_MAX_PEERS) PT.hiz[tid] = peer_args.hiz[tid / DFX_MAX_PEERS][tid % DFX_MAX_PEERS];
        if (tid < DFX_MAX_PEERS) PT.color[tid] = peer_args.color[tid], PT.normal[tid] = peer_args.normal[tid];
        {
            const int row = tid << kPeerBlockShift;
            int       o   = 0;
            for (int r = 1; r < peer_args.count; ++r) o = row >= peer_args.row_begin[r] ? r : o;
            PT.owner[tid] = (uint8_t)o;
        }
    }
    __syncthreads();
    const CamS& cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_rad.w || y >= y1) return;
    if (!__ldg(&mask.at(x, y)))
    {
        // cleared targets (ScreenSpaceReflection.cpp:991-994)
        st_cs(&out_rad.at(x, y), make_float4(0.f, 0.f, 0.f, 0.f));
        st_cs(&out_dir.at(x, y), make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    const float sw = cam.vw, sh = cam.vh;
    // FEATURE_FLAG_HALF_RESOLUTION: the targets are W/2 x H/2 and each target pixel traces the ray of one of its four
    // full-resolution pixels, picked by a 4x4 pattern of 2-bit offsets (PostFX_Common.fxh:45-55; :283-288)
    int px = x, py = y;
    if (half)
    {
        const unsigned idx = (1320229860u >> ((((unsigned)x & 3u) << 3) + (((unsigned)y & 3u) << 1))) & 3u;
        px = 2 * x + (int)(idx & 1u), py = 2 * y + (int)(idx >> 1);
    }
    const float u = (float(px) + 0.5f) * cam.ivw, v = (float(py) + 0.5f) * cam.ivh;
    const float3 nws = xyz(normal.ld(px, py));
    const float3 nvs = mul_dir(nws, S.view);
    const float  rough = __ldg(&roughness.at(px, py));

    const bool  mirror  = rough < 0.01f;
    const int   baseMip = mirror ? 0 : (int)A.MostDetailedMip;
    const float inv0    = 1.0f / float(1 << baseMip);
    const float mrx = sw * inv0, mry = sh * inv0;

    const float  originD = hiz_load(hiz, (int)(u * mrx), (int)(v * mry), baseMip);
    const float3 originVS = screen_to_view(u, v, originD, cam);

    // SampleReflectionVector :254-278
    float3 dirVS;
    float  pdf;
    {
        const float3 V = -fnormalize(originVS);
        const float  a = rough * rough;
        const float3 N = nvs;
        const float3 T = fnormalize(cross(N, fabsf(N.y) > 0.5f ? make_float3(1.f, 0.f, 0.f) : make_float3(0.f, 1.f, 0.f)));
        const float3 B = cross(T, N);
        float2       xi = __ldg(&noise.at(x & 127, y & 127));
        xi.y            = lerpf(xi.y, 0.0f, A.GGXImportanceSampleBias);
        const float3 Vts = make_float3(dot(T, V), dot(B, V), dot(N, V));
        // SmithGGXSampleVisibleNormalSC (PBR_Common.fxh:278-296) with ax == ay == a
        const float3 Vs  = fnormalize(make_float3(Vts.x * a, Vts.y * a, Vts.z));
        const float  phi = 2.0f * 3.141592653589793f * xi.x;
        const float  Z   = (1.0f - xi.y) * (1.0f + Vs.z) - Vs.z;
        const float  st  = fsqrt(fminf(fmaxf(1.0f - Z * Z, 0.0f), 1.0f));
        float        sp, cp;
        __sincosf(phi, &sp, &cp);
        const float3 Hh = make_float3(st * cp, st * sp, Z) + Vs;
        const float3 Hm = fnormalize(make_float3(a * Hh.x, a * Hh.y, Hh.z));
        // reflect(-Vts, Hm) = -Vts - 2*dot(Hm, -Vts)*Hm
        const float3 I  = -Vts;
        const float3 Lts = I - 2.0f * dot(Hm, I) * Hm;
        const float  D  = ndf_ggx(Hm.z, a);
        const float  G1 = smith_ggx_masking(Vts.z, a);
        pdf             = G1 * D / (4.0f * Vts.z + kFltEps);
        dirVS           = Lts.x * T + Lts.y * B + Lts.z * N;
    }
    // ProjectDirection (PostFX_Common.fxh:94-97)
    const float3 endSS = project_position(originVS + dirVS, S.proj);
    const float3 O     = make_float3(u, v, originD);
    const float3 Dr    = endSS - O;
    const float3 dirWS = mul_dir(dirVS, S.view_inv);

    // HierarchicalRaymarch :139-189
    float3 pos;
    bool   validHit;
    {
        const float3 invD = make_float3(Dr.x != 0.0f ? 1.0f / Dr.x : kFltMax, Dr.y != 0.0f ? 1.0f / Dr.y : kFltMax, Dr.z != 0.0f ? 1.0f / Dr.z : kFltMax);
        int   mip = baseMip;
        const float resx = sw * inv0, resy = sh * inv0;
        const float irx = 1.0f / resx, iry = 1.0f / resy;
        float uox = 0.005f * float(1 << baseMip) / sw, uoy = 0.005f * float(1 << baseMip) / sh;
        uox = Dr.x < 0.0f ? -uox : uox, uoy = Dr.y < 0.0f ? -uoy : uoy;
        const float fox = Dr.x < 0.0f ? 0.0f : 1.0f, foy = Dr.y < 0.0f ? 0.0f : 1.0f;

        float t;
        {
            // InitialAdvanceRay :66-86
            float px = (floorf(resx * O.x) + fox) * irx + uox, py = (floorf(resy * O.y) + foy) * iry + uoy;
            float tx = px * invD.x - O.x * invD.x, ty = py * invD.y - O.y * invD.y;
            t   = fminf(tx, ty);
            pos = O + t * Dr;
        }
        // The loop is issue-bound (one dependent Hi-Z load per iteration, ~36 iterations per ray): per-level constants come
        // from one 32-byte shared-memory entry, the level update is branch-free, the trip counter counts down.
        const float oxi = O.x * invD.x, oyi = O.y * invD.y, ozi = O.z * invD.z;
        int         left = (int)min(A.MaxTraversalIntersections, 0x7fffffffu);
#if DFX_INTERSECT_V2
        // Opt-in (-DDFX_INTERSECT_V2=1), same ray paths bit for bit, fewer instructions per step: the depth-plane crossing is
        // switched off through its loop-invariant coefficients (surf * 0 + FLT_MAX is FLT_MAX exactly) instead of a select
        // per step, and the level moves by min(mip + 1, 6) / mip - 1.
        const bool  zOn = REV ? Dr.z < 0.0f : Dr.z > 0.0f;
        const float rz = zOn ? invD.z : 0.0f, oz = zOn ? ozi : -kFltMax;
        while (left > 0 && mip >= baseMip)
        {
            const HizLevel L  = lvl[mip];
            const float    mx = L.resx * pos.x, my = L.resy * pos.y;
            const float    surf = hiz_load(L, (int)mx, (int)my); // sharded frames too: the pyramid is complete on every GPU (gathered before the march)
            const float px = (floorf(mx) + fox) * L.irx + uox, py = (floorf(my) + foy) * L.iry + uoy;
            const float tx = px * invD.x - oxi, ty = py * invD.y - oyi;
            const float tz = surf * rz - oz;
            const float tmin  = fminf(fminf(tx, ty), tz);
            const bool  above = REV ? surf < pos.z : surf > pos.z;
            const bool  skipped = (__float_as_uint(tmin) != __float_as_uint(tz)) && above;
            t   = above ? tmin : t;
            pos = O + t * Dr;
            mip = skipped ? min(mip + 1, 6) : mip - 1;
            --left;
        }
#else
        while (left > 0 && mip >= baseMip)
        {
            const HizLevel L  = lvl[mip];
            const float    mx = L.resx * pos.x, my = L.resy * pos.y;
            const float    surf = hiz_load(L, (int)mx, (int)my);
            // AdvanceRay :88-136
            const float px = (floorf(mx) + fox) * L.irx + uox, py = (floorf(my) + foy) * L.iry + uoy;
            const float tx = px * invD.x - oxi, ty = py * invD.y - oyi;
            const float tz = (REV ? Dr.z < 0.0f : Dr.z > 0.0f) ? surf * invD.z - ozi : kFltMax; // :109-113
            const float tmin  = fminf(fminf(tx, ty), tz);
            const bool  above = REV ? surf < pos.z : surf > pos.z;                             // :118-124
            const bool  skipped = (__float_as_uint(tmin) != __float_as_uint(tz)) && above;
            t   = above ? tmin : t;
            pos = O + t * Dr;
            mip += skipped ? (mip >= 6 ? 0 : 1) : -1; // a skip at the coarsest level stays there (:179-183)
            --left;
        }
#endif
        validHit = true; // i <= MaxTraversalIntersections always holds at loop exit (:187)
    }
    const float3 hitVS = screen_to_view(pos.x, pos.y, pos.z, cam);

    float hpx = pos.x, hpy = pos.y; // previous-frame hit position
    if (PREV_FRAME)
    {
        float2 mv = load0(motion, (int)(sw * pos.x), (int)(sh * pos.y));
        hpx = pos.x - mv.x * 0.5f, hpy = pos.y - mv.y * -0.5f;
    }

    // ValidateHit :201-242
    float  confidence = 0.0f;
    float4 peer_colour = make_float4(0.f, 0.f, 0.f, 0.f);
    if (validHit && !(pos.x < 0.0f || pos.y < 0.0f || pos.x > 1.0f || pos.y > 1.0f))
    {
        const float mdx = fabsf(pos.x - u), mdy = fabsf(pos.y - v);
        if (!(mdx < (2.0f / sw) && mdy < (2.0f / sh)))
        {
            const int   tx = (int)(sw * pos.x), ty = (int)(sh * pos.y);
            const float surfD = hiz_load(lvl[0], tx, ty);
            if (!is_background(surfD, REV))
            {
                const float3 hitN = xyz(hit_load0<true>(normal, PT, tx, ty));
                // sharded frame: both texels at the hit may live on another GPU. The colour is requested together with the normal (it is
                // only needed if the hit survives the tests below) so that a ray pays ONE trip over the link, not two in a chain.
                if (PEER) peer_colour = hit_load0<false>(color, PT, tx, ty);
                if (!(dot(hitN, dirWS) > 0.0f))
                {
                    const float3 surfVS = screen_to_view(pos.x, pos.y, surfD, cam);
                    const float3 dd     = surfVS - hitVS;
                    const float  dist   = length(dd);
                    float        vig    = edge_vignette(pos.x, pos.y, sw, sh);
                    if (PREV_FRAME) vig = fminf(edge_vignette(hpx, hpy, sw, sh), vig);
                    float c = 1.0f - smoothstepf(0.0f, A.DepthBufferThickness, dist * (1.0f / (surfVS.z + kFltEps)));
                    c *= c;
                    confidence = vig * c;
                }
            }
        }
    }
    float3 radiance = make_float3(0.f, 0.f, 0.f);
    if (confidence > 0.0f) radiance = PEER ? xyz(peer_colour) : xyz(hit_load0<false>(color, PT, (int)(sw * hpx), (int)(sh * hpy))); // PEER implies !PREV_FRAME: same texel
    const float3 dv = hitVS - originVS;
    st_cs(&out_rad.at(x, y), f4(radiance, confidence));
    st_cs(&out_dir.at(x, y), f4(dirWS * length(dv), pdf));
}

// ---------------------------------------------------------------------------------------------------------------------
// S5: spatial reconstruction — SSR_ComputeSpatialReconstruction.fx:114-172
// ---------------------------------------------------------------------------------------------------------------------
__constant__ float3 kSsrPoisson8[8] = {{-0.4706069f, -0.4427112f, +0.6461146f}, {-0.9057375f, +0.3003471f, +0.9542373f},
                                       {-0.3487388f, +0.4037880f, +0.5335386f}, {+0.1023042f, +0.6439373f, +0.6520134f},
                                       {+0.5699277f, +0.3513750f, +0.6695386f}, {+0.2939128f, -0.1131226f, +0.3149309f},
                                       {+0.7836658f, -0.4208784f, +0.8895339f}, {+0.1564120f, -0.8198990f, +0.8346850f}};
__constant__ float kSsrPoisson8Weight[8] = {0.77283178f, 0.570022457f, 0.838854363f, 0.769187387f, 0.758268871f, 0.940613337f, 0.613583686f, 0.650469323f};
struct SpatialCam
{
    CamS c;
    Mat4 vp_inv;
};

template <bool N16>
__global__ void __launch_bounds__(256, DFX_OCC_SSR_SPATIAL) ssr_spatial_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssr_attribs A,
                                                          View<const float> roughness, View<const uint8_t> mask, TexRGBA<N16> normal,
                                                          View<const float> depth, View<const float4> raydir, View<const float4> radiance,
                                                          View<float4> out_rad, View<float> out_var, View<float> out_depth, int y0, int y1, int half)
{
    __shared__ SpatialCam S;
    if (threadIdx.x == 0 && threadIdx.y == 0) load_cam(S.c, &cams[0]), load_mat(S.vp_inv, cams[0].mViewProjInv);
    __syncthreads();
    const CamS& cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_rad.w || y >= y1) return;
    if (!__ldg(&mask.at(x, y))) return;

    const int    W = (int)cam.vw, H = (int)cam.vh;
    const float  posx = float(x) + 0.5f, posy = float(y) + 0.5f;
    const float3 camPos = make_float3(cam.px, cam.py, cam.pz);
    const float3 pws = inv_project_position(posx * cam.ivw, posy * cam.ivh, __ldg(&depth.at(x, y)), S.vp_inv);
    const float3 nws = xyz(normal.ld(x, y));
    const float3 toCam = camPos - pws;
    const float  camDist2 = dot(toCam, toCam), invCamDist = frsqrt(camDist2);
    const float3 vws = toCam * invCamDist;
    const float  NdotV = saturate(dot(nws, vws));
    const float  rough = __ldg(&roughness.at(x, y));
    const float  radius = lerpf(0.0f, A.SpatialReconstructionRadius, saturate(5.0f * rough));
    float        rs, rc;
    __sincosf(2.0f * kPi * bayer4x4((uint32_t)x, (uint32_t)y, cam.frame_index), &rs, &rc);
    // per-pixel invariants of SmithGGXVisibilityCorrelated / NormalDistribution_GGX (PBR_Common.fxh:107-124, :181-195)
    const float a    = rough * rough, a2 = a * a;
    const float ad   = fmaxf(a, 1e-3f), ad2 = ad * ad;
    const float visV = NdotV * NdotV * (1.0f - a2) + a2; // under the sqrt of GGXV

    const float sscale = half ? 0.5f : 1.0f, sbias = half ? 0.5f : 0.0f, sbx = half ? float(x) : posx, sby = half ? float(y) : posy;
    const int   smaxx = (half ? (int)(0.5f * cam.vw) : W) - 1, smaxy = (half ? (int)(0.5f * cam.vh) : H) - 1;
    float4 colorSum = make_float4(0.f, 0.f, 0.f, 0.f);
    float  wsum = 0.0f, variance = 0.0f, mean = 0.0f, nearest = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const float3 P  = kSsrPoisson8[i];
        const float  xi = P.x * rc + P.y * rs, yi = P.x * -rs + P.y * rc;
        // full resolution: int2(Position + Radius * Xi); half-size intersect targets: int2(0.5 * (floor(Position) + Radius * Xi) + 0.5)
        // (:153-157). One expression for both — the scale 1 and the bias 0 of the full-resolution case are exact.
        const int sx = min(max((int)(sscale * (sbx + radius * xi) + sbias), 0), smaxx), sy = min(max((int)(sscale * (sby + radius * yi) + sbias), 0), smaxy);
        const float  ws = kSsrPoisson8Weight[i]; // exp(-z^2 / (2 * 0.9^2)), a constant per disk sample
        // ComputeWeightRayLength :60-86
        float        weight, raylen;
        const float4 rd = __ldg(&raydir.at(sx, sy));
        const float  len2 = dot(xyz(rd), xyz(rd)), invLen = frsqrt(len2), len = len2 * invLen;
        if (!(len >= 1e-6f)) // also catches len2 == 0 (0 * inf = NaN)
        {
            weight = 1e-6f, raylen = 1e-6f;
        }
        else
        {
            const float3 L  = xyz(rd) * invLen;
            const float3 Hh = fnormalize(L + vws);
            const float  NdotH = saturate(dot(nws, Hh)), NdotL = saturate(dot(nws, L));
            const float  ggxv = NdotL * fsqrt(fmaxf(visV, 1e-7f));
            const float  ggxl = NdotV * fsqrt(fmaxf(NdotL * NdotL * (1.0f - a2) + a2, 1e-7f));
            const float  f    = NdotH * NdotH * ad2 + (1.0f - NdotH * NdotH);
            // Vis * D * NdotL * ws / max(pdf, 1e-5)
            const float  num = 0.5f * ad2 * NdotL * ws;
            const float  den = (ggxv + ggxl) * fmaxf(3.141592653589793f * f * f, 1e-9f) * fmaxf(rd.w, 1e-5f);
            weight = fmaxf(fdiv(num, den), 1e-6f);
            raylen = len;
        }
        const float4 c = __ldg(&radiance.at(sx, sy));
        // ComputeWeightedVariance :90-100
        colorSum = colorSum + weight * c;
        wsum += weight;
        const float value = luminance(xyz(c)), prevMean = mean;
        mean += weight * frcp(wsum) * (value - prevMean);
        variance += weight * (value - prevMean) * (value - mean);
        if (weight > 1.0e-6f) nearest = fmaxf(raylen, nearest);
    }
    const float iden = frcp(fmaxf(wsum, 1e-6f));
    out_rad.at(x, y) = colorSum * iden;
    out_var.at(x, y) = variance * iden;
    out_depth.at(x, y) = camz_to_depth(camDist2 * invCamDist + nearest, cam);
}

// ---------------------------------------------------------------------------------------------------------------------
// S6: temporal accumulation — SSR_ComputeTemporalAccumulation.fx:224-263
// ---------------------------------------------------------------------------------------------------------------------
struct SsrTemporalCam
{
    CamS c, p;
    Mat4 curr_vp_inv, prev_vp;
    float pjx, pjy;
};

// ComputeDisocclusion :111-116 is exp(-r) with r = |cz - pz| / max(|cz|, |pz|, 1e-6); it is only ever compared with a
// threshold T, and exp(-r) > T  <=>  r < -ln(T), so the kernels evaluate r and compare against -ln(T).
DFX_HD float disocclusion_ratio(float cz, float pz)
{
    cz = fabsf(cz), pz = fabsf(pz);
    return fdiv(fabsf(cz - pz), fmaxf(fmaxf(cz, pz), 1e-6f));
}
constexpr float kNegLn090 = 0.105360516f; // -ln(SSR_DISOCCLUSION_THRESHOLD)
constexpr float kNegLn045 = 0.798507696f; // -ln(SSR_DISOCCLUSION_THRESHOLD / 2)

// PEER: the three previous-frame planes (read at the reprojected position, which no fixed halo bounds) are loaded from the GPU that
// owns the row (PeerView); everything else is read at the pixel or within +-1 row of it.
struct NoPeerMap
{
};
template <bool PEER, bool M16>
__global__ void __launch_bounds__(256, DFX_OCC_SSR_TEMPORAL) ssr_temporal_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssr_attribs A,
                                                           View<const uint8_t> mask, TexRG<M16> motion, View<const float> hit_depth,
                                                           View<const float> curr_depth, View<const float4> curr_rad, View<const float> curr_var,
                                                           View<const float> prev_depth_, View<const float4> prev_rad_, View<const float> prev_var_,
                                                           View<float4> out_rad, View<float> out_var, int y0, int y1,
                                                           const __grid_constant__ typename std::conditional<PEER, PeerMap, NoPeerMap>::type peer_map)
{
    typename std::conditional<PEER, PeerView<float>, View<const float>>::type   prev_depth, prev_var;
    typename std::conditional<PEER, PeerView<float4>, View<const float4>>::type prev_rad;
    if constexpr (PEER)
    {
        prev_depth = PeerView<float>{prev_depth_.p, prev_depth_.pitch, prev_depth_.w, prev_depth_.h, &peer_map};
        prev_var   = PeerView<float>{prev_var_.p, prev_var_.pitch, prev_var_.w, prev_var_.h, &peer_map};
        prev_rad   = PeerView<float4>{prev_rad_.p, prev_rad_.pitch, prev_rad_.w, prev_rad_.h, &peer_map};
    }
    else
        prev_depth = prev_depth_, prev_var = prev_var_, prev_rad = prev_rad_;
    __shared__ SsrTemporalCam S;
    if (threadIdx.x == 0 && threadIdx.y == 0)
    {
        load_cam(S.c, &cams[0]), load_cam(S.p, &cams[1]);
        load_mat(S.curr_vp_inv, cams[0].mViewProjInv), load_mat(S.prev_vp, cams[1].mViewProj);
        S.pjx = cams[1].f2Jitter[0], S.pjy = cams[1].f2Jitter[1];
    }
    __syncthreads();
    const CamS& cam = S.c;
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out_rad.w || y >= y1) return;
    if (!__ldg(&mask.at(x, y))) return;

    const int   W = (int)cam.vw, H = (int)cam.vh;
    const float posx = float(x) + 0.5f, posy = float(y) + 0.5f;

    // ComputePixelStatistic :119-145
    float4 m1 = make_float4(0.f, 0.f, 0.f, 0.f), m2 = m1;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
        {
            const float4 s = __ldg(&curr_rad.at(min(max(x + dx, 0), W - 1), min(max(y + dy, 0), H - 1)));
            m1 = m1 + s, m2 = m2 + s * s;
        }
    const float4 mean = m1 * (1.0f / 9.0f);
    const float4 var  = (m2 * (1.0f / 9.0f)) - (mean * mean);
    const float4 sd   = make_float4(fsqrt(fmaxf(var.x, 0.f)), fsqrt(fmaxf(var.y, 0.f)), fsqrt(fmaxf(var.z, 0.f)), fsqrt(fmaxf(var.w, 0.f)));

    const float depth = __ldg(&curr_depth.at(x, y));
    const float hitD  = __ldg(&hit_depth.at(x, y));
    float2      mv    = motion.ld(x, y);
    mv.x *= 0.5f, mv.y *= -0.5f;

    const float ipx = posx - mv.x * cam.vw, ipy = posy - mv.y * cam.vh; // PrevIncidentPoint
    // ComputeReflectionHitPosition :102-108
    float rhx, rhy;
    {
        const float  tu = posx * cam.ivw + 0.5f * cam.jx, tv = posy * cam.ivh + -0.5f * cam.jy;
        const float3 pws = inv_project_position(tu, tv, hitD, S.curr_vp_inv);
        const float3 puv = project_position(pws, S.prev_vp);
        rhx = (puv.x - 0.5f * S.pjx) * cam.vw, rhy = (puv.y - -0.5f * S.pjy) * cam.vh;
    }
    const float4 cI = sample_linear_clamp(prev_rad, ipx * cam.ivw, ipy * cam.ivh);
    const float4 cR = sample_linear_clamp(prev_rad, rhx * cam.ivw, rhy * cam.ivh);
    const float  lm = luminance(xyz(mean));
    const float  dI = fabsf(luminance(xyz(cI)) - lm), dR = fabsf(luminance(xyz(cR)) - lm);
    const float  ppx = dI < dR ? ipx : rhx, ppy = dI < dR ? ipy : rhy;

    // ComputeReprojection :147-221
    const float currZ = depth_to_camz(depth, cam);
    float4      rcol;
    float       rpx = ppx, rpy = ppy;
    bool        ok;
    {
        const float pz = depth_to_camz(load0(prev_depth, (int)ppx, (int)ppy), S.p);
        rcol           = sample_linear_clamp(prev_rad, ppx * cam.ivw, ppy * cam.ivh);
        ok             = disocclusion_ratio(currZ, pz) < kNegLn090;
    }
    if (!ok)
    {
        float bw00 = 0.f, bw10 = 0.f, bw01 = 0.f, bw11 = 0.f, best = 0.0f;
        int   bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;
        bool  done = false;
        for (int dy = -1; dy <= 1 && !done; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
            {
                const float lx = ppx + float(dx), ly = ppy + float(dy);
                const Bilin b  = bilinear_uc(lx, ly, curr_depth.w, curr_depth.h);
                auto pass = [&](int sx, int sy) { return disocclusion_ratio(currZ, depth_to_camz(load0(prev_depth, sx, sy), S.p)) < kNegLn045 ? 1.0f : 0.0f; };
                const float w00 = b.w00 * pass(b.x0, b.y0), w10 = b.w10 * pass(b.x1, b.y0), w01 = b.w01 * pass(b.x0, b.y1), w11 = b.w11 * pass(b.x1, b.y1);
                const float tot = w00 * 1.0f + w10 * 1.0f + w01 * 1.0f + w11 * 1.0f;
                if (tot > best)
                {
                    best = tot, bw00 = w00, bw10 = w10, bw01 = w01, bw11 = w11;
                    bx0 = b.x0, by0 = b.y0, bx1 = b.x1, by1 = b.y1;
                    rpx = lx, rpy = ly;
                    if (best > 0.9f)
                    {
                        done = true;
                        break;
                    }
                }
            }
        ok = best > 0.1f;
        if (ok)
            rcol = (load0(prev_rad, bx0, by0) * bw00 + load0(prev_rad, bx1, by0) * bw10 + load0(prev_rad, bx0, by1) * bw01 + load0(prev_rad, bx1, by1) * bw11) / best;
    }
    ok = ok && (rpx >= 0.0f && rpy >= 0.0f && rpx < cam.vw && rpy < cam.vh);

    const float4 cr = __ldg(&curr_rad.at(x, y));
    if (ok)
    {
        const float4 lo = mean - 2.5f * sd, hi = mean + 2.5f * sd;
        const float4 pr = make_float4(fminf(fmaxf(rcol.x, lo.x), hi.x), fminf(fmaxf(rcol.y, lo.y), hi.y), fminf(fmaxf(rcol.z, lo.z), hi.z),
                                      fminf(fmaxf(rcol.w, lo.w), hi.w));
        const float  pv = sample_linear_clamp(prev_var, rpx * cam.ivw, rpy * cam.ivh);
        out_rad.at(x, y) = lerp4(cr, pr, A.TemporalRadianceStabilityFactor);
        out_var.at(x, y) = lerpf(__ldg(&curr_var.at(x, y)), pv, A.TemporalVarianceStabilityFactor);
    }
    else
    {
        out_rad.at(x, y) = cr;
        out_var.at(x, y) = 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// S7: bilateral cleanup — SSR_ComputeBilateralCleanup.fx:49-97. ddx/ddy(CameraZ) are the 2x2 pixel-quad finite
// differences v(x|1) - v(x&~1), v(y|1) - v(y&~1) (coordinates clamped at odd image edges).
// ---------------------------------------------------------------------------------------------------------------------
template <bool N16>
__global__ void __launch_bounds__(256) ssr_bilateral_kernel(const dfx_camera_attribs* __restrict__ cams, dfx_ssr_attribs A,
                                                            View<const uint8_t> mask, View<const float> depth, TexRGBA<N16> normal,
                                                            View<const float> roughness, View<const float4> radiance, View<const float> variance,
                                                            View<float4> out, int y0, int y1, int rev)
{
    __shared__ CamS cam;
    if (threadIdx.x == 0 && threadIdx.y == 0) load_cam(cam, &cams[0]);
    __syncthreads();
    const PixelXY pix = cta_pixel(y0);
    const int     x = pix.x, y = pix.y;
    if (x >= out.w || y >= y1) return;
    if (!__ldg(&mask.at(x, y)))
    {
        st_cs(&out.at(x, y), make_float4(0.f, 0.f, 0.f, 0.f)); // cleared target (…cpp:1097-1099)
        return;
    }
    const int   W = (int)cam.vw, H = (int)cam.vh;
    const float rough = __ldg(&roughness.at(x, y));
    const float var   = __ldg(&variance.at(x, y));
    const float3 nws  = xyz(normal.ld(x, y));
    // The depth edge-stopping weight exp(-|dz| / (|grad . d| + 1e-6)) divides by a quantity that is ~0 on flat surfaces, so it
    // amplifies the last bits of the camera-space Z: this pass keeps the correctly-rounded division for Z.
    auto         camz_precise = [&](float dpt) { return (cam.m32 - dpt * cam.m33) / (dpt * cam.m23 - cam.m22); };
    const float target = saturate(8.0f * rough);
    const float radius = lerpf(0.0f, var > 0.001f ? 2.0f : 0.0f, target);
    const float sigma  = A.BilateralCleanupSpatialSigmaFactor;
    const int   er     = (int)fminf(2.0f * sigma, radius);
    float4      result = __ldg(&radiance.at(x, y));
    if (var > 0.00005f && er > 0)
    {
        // camera Z and its quad derivatives are only needed by the filter branch. On an odd-sized target the quad partner of
        // the last column / row lies outside it: Direct3D shades it as a helper lane whose Load returns 0 (load0, not a clamp)
        const float  camZ = camz_precise(__ldg(&depth.at(x, y)));
        auto         cz   = [&](int sx, int sy) { return camz_precise(load0(depth, sx, sy)); };
        const float  gx = cz(x | 1, y) - cz(x & ~1, y), gy = cz(x, y | 1) - cz(x, y & ~1);
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        float  wsum = 0.0f;
        const float inv_sigma2 = frcp(sigma * sigma);
        for (int dx = -er; dx <= er; ++dx)
            for (int dy = -er; dy <= er; ++dy)
            {
                const int   sx = min(max(x + dx, 0), W - 1), sy = min(max(y + dy, 0), H - 1);
                const float sd = __ldg(&depth.at(sx, sy)), sr = __ldg(&roughness.at(sx, sy));
                if (is_reflection_sample(sr, sd, A.RoughnessThreshold, rev))
                {
                    const float4 srad = __ldg(&radiance.at(sx, sy));
                    const float3 sn   = xyz(normal.ld(sx, sy));
                    const float  sz   = camz_precise(sd);
                    const float  fx = float(dx), fy = float(dy);
                    const float  ws = __expf(-0.5f * (fx * fx + fy * fy) * inv_sigma2);
                    const float  wz = __expf(-fabsf(camZ - sz) / (fabsf(fx * gx + fy * gy) + 1e-6f));
                    // x^128 by seven squarings (exact to a few ulp, no log/exp round trip)
                    float        wn = fmaxf(0.0f, dot(nws, sn));
                    wn *= wn, wn *= wn, wn *= wn, wn *= wn, wn *= wn, wn *= wn, wn *= wn;
                    const float  w  = ws * wn * wz;
                    wsum += w;
                    csum = csum + w * srad;
                }
            }
        result = csum / fmaxf(wsum, 1.0e-6f);
    }
    st_cs(&out.at(x, y), make_float4(result.x, result.y, result.z, result.w * A.AlphaInterpolation));
}

static bool make_hiz(const dfx_pyramid* p, HizView& v)
{
    if (!p || p->levels < 1 || p->levels > DFX_MAX_MIPS) return false;
    v.levels = p->levels;
    for (int i = 0; i < p->levels; ++i)
    {
        if (!make_view<const float>(&p->level[i], DFX_FORMAT_R32F, v.lv[i])) return false;
        if (i > 0 && (v.lv[i].w != max(v.lv[0].w >> i, 1) || v.lv[i].h != max(v.lv[0].h >> i, 1))) return false;
    }
    return true;
}

} // namespace dfx

using namespace dfx;

namespace dfx
{
void preload_postfx_kernels();
// Everything a strip-sharded SSR frame launches, loaded up front: a kernel that is loaded lazily while a flag-wait kernel spins can
// deadlock ranks that share a process (module loading may synchronise the device).
void preload_ssr_strip_kernels()
{
    preload_postfx_kernels();
    cudaFuncAttributes fa;
    (void)cudaFuncGetAttributes(&fa, pyramid_level_kernel<HizOp>);
    (void)cudaFuncGetAttributes(&fa, pyramid_tail_kernel<HizOp>);
    (void)cudaFuncGetAttributes(&fa, pyramid_tile_kernel<HizOp, true>);
    (void)cudaFuncGetAttributes(&fa, pyramid_tile_kernel<HizOp, false>);
    (void)cudaFuncGetAttributes(&fa, ssr_mask_kernel<DFX_FORMAT_RGBA32F>);
    (void)cudaFuncGetAttributes(&fa, ssr_intersect_kernel<false, true, false, false>);
    (void)cudaFuncGetAttributes(&fa, ssr_intersect_kernel<false, true, true, false>);
    (void)cudaFuncGetAttributes(&fa, ssr_intersect_kernel<false, false, false, false>);
    (void)cudaFuncGetAttributes(&fa, ssr_intersect_kernel<false, false, true, false>);
    (void)cudaFuncGetAttributes(&fa, ssr_spatial_kernel<false>);
    (void)cudaFuncGetAttributes(&fa, ssr_temporal_kernel<true, false>);
    (void)cudaFuncGetAttributes(&fa, ssr_temporal_kernel<false, false>);
    (void)cudaFuncGetAttributes(&fa, ssr_bilateral_kernel<false>);
    (void)cudaGetLastError();
}
} // namespace dfx

extern "C" dfx_status dfx_pass_ssr_hiz(void* stream, const dfx_pyramid* pyr, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssr_hiz");
    DFX_REQUIRE(pyr && pyr->levels >= 1 && pyr->levels <= DFX_MAX_MIPS, "bad Hi-Z pyramid");
    PyrPlanes<1> P;
    P.levels = pyr->levels;
    for (int i = 0; i < pyr->levels; ++i)
    {
        DFX_REQUIRE(make_view<float>(&pyr->level[i], DFX_FORMAT_R32F, P.lv[0][i]), "bad Hi-Z level %d", i);
        DFX_REQUIRE(i == 0 || (P.lv[0][i].w == max(P.lv[0][0].w >> i, 1) && P.lv[0][i].h == max(P.lv[0][0].h >> i, 1)), "Hi-Z level %d has the wrong size", i);
    }
    const int H = P.lv[0][0].h;
    DFX_REQUIRE(rows_ok(rows, H) && rows.y0 % 64 == 0 && (rows.y1 % 64 == 0 || rows.y1 == H), "pyramid passes need 64-row aligned strips");
    return build_pyramid(stream, HizOp{reversed_depth(&pyr->level[0])}, P, 6, rows, "ssr_hiz pyramid kernel");
}

#define DFX_GRID(w, rows) dim3 block(32, 8), grid(div_up(w, 32), div_up(rows.y1 - rows.y0, 8))

extern "C" dfx_status dfx_pass_ssr_mask_roughness(void* stream, const dfx_ssr_attribs* attribs, const dfx_plane* material, const dfx_plane* depth,
                                                  const dfx_plane* roughness, const dfx_plane* mask, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssr_mask_roughness");
    DFX_REQUIRE(attribs, "null argument");
    Tex4 m;
    DFX_REQUIRE(make_tex4(material, m, true), "bad material plane (RGBA32F, RGBA16F or RG8U)");
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    DFX_VIEW(float, r, roughness, DFX_FORMAT_R32F);
    DFX_VIEW(uint8_t, k, mask, DFX_FORMAT_R8U);
    DFX_SAME_SIZE(d, m);
    DFX_SAME_SIZE(d, r);
    DFX_SAME_SIZE(d, k);
    DFX_REQUIRE(rows_ok(rows, d.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(d.w, rows);
    if (m.fmt == DFX_FORMAT_RGBA32F) ssr_mask_kernel<DFX_FORMAT_RGBA32F><<<grid, block, 0, as_stream(stream)>>>(*attribs, m, d, r, k, rows.y0, rows.y1, reversed_depth(depth));
    else if (m.fmt == DFX_FORMAT_RGBA16F) ssr_mask_kernel<DFX_FORMAT_RGBA16F><<<grid, block, 0, as_stream(stream)>>>(*attribs, m, d, r, k, rows.y0, rows.y1, reversed_depth(depth));
    else ssr_mask_kernel<DFX_FORMAT_RG8U><<<grid, block, 0, as_stream(stream)>>>(*attribs, m, d, r, k, rows.y0, rows.y1, reversed_depth(depth));
    DFX_LAUNCHED("ssr_mask_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssr_downsample_mask(void* stream, const dfx_ssr_attribs* attribs, const dfx_plane* roughness, const dfx_plane* depth,
                                                   const dfx_plane* mask_half, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssr_downsample_mask");
    DFX_REQUIRE(attribs, "null argument");
    DFX_VIEW(const float, r, roughness, DFX_FORMAT_R32F);
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    DFX_VIEW(uint8_t, k, mask_half, DFX_FORMAT_R8U);
    DFX_SAME_SIZE(d, r);
    DFX_REQUIRE(k.w == d.w / 2 && k.h == d.h / 2, "the downsampled mask must be width/2 x height/2 of the depth plane");
    DFX_REQUIRE(rows_ok(rows, k.h), "bad row range (rows of the half-size mask)");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(k.w, rows);
    ssr_downsample_mask_kernel<<<grid, block, 0, as_stream(stream)>>>(*attribs, r, d, k, rows.y0, rows.y1, reversed_depth(depth));
    DFX_LAUNCHED("ssr_downsample_mask_kernel");
    return DFX_OK;
}

static dfx_status ssr_intersect_impl(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs, uint32_t flags,
                                     const dfx_peer_set* peers, const dfx_plane* color, const dfx_plane* normal, const dfx_plane* roughness,
                                     const dfx_plane* mask, const dfx_plane* blue_noise_xy, const dfx_pyramid* hiz, const dfx_plane* motion,
                                     const dfx_plane* out_radiance, const dfx_plane* out_raydir_pdf, dfx_rows rows)
{
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_TEX4(c, color);
    DFX_TEX4(n, normal);
    DFX_VIEW(const float, r, roughness, DFX_FORMAT_R32F);
    DFX_VIEW(const uint8_t, k, mask, DFX_FORMAT_R8U);
    DFX_VIEW(const float2, bn, blue_noise_xy, DFX_FORMAT_RG32F);
    DFX_VIEW(float4, orad, out_radiance, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, odir, out_raydir_pdf, DFX_FORMAT_RGBA32F);
    HizView H;
    DFX_REQUIRE(make_hiz(hiz, H), "bad Hi-Z pyramid");
    const bool rev = reversed_depth(&hiz->level[0]) != 0; // level 0 is the depth buffer
    DFX_SAME_SIZE(c, n);
    DFX_SAME_SIZE(c, r);
    // Targets (and mask) of half the colour plane's size mean FEATURE_FLAG_HALF_RESOLUTION (…cpp:201-213: Radiance and
    // RayDirectionPDF are W/2 x H/2; the mask is then the downsampled one of dfx_pass_ssr_downsample_mask)
    const int half = (orad.w != c.w || orad.h != c.h) ? 1 : 0;
    DFX_REQUIRE(!half || (orad.w == c.w / 2 && orad.h == c.h / 2), "the intersect targets must have the size of the colour plane or half of it");
    DFX_REQUIRE(!half || !peers, "half-resolution SSR is not supported on peer-sharded frames");
    DFX_SAME_SIZE(orad, k);
    DFX_SAME_SIZE(orad, odir);
    DFX_SAME_SIZE(c, H.lv[0]);
    DFX_REQUIRE(bn.w == 128 && bn.h == 128, "blue noise must be 128x128");
    DFX_REQUIRE(rows_ok(rows, orad.h), "bad row range (rows of the intersect targets)");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(orad.w, rows);
    const bool g16 = is16(c);
    DFX_REQUIRE(is16(n) == g16, "colour and normal must both be RGBA16F or both RGBA32F");
    if (peers)
    {
        DFX_REQUIRE((flags & DFX_SSR_FEATURE_FLAG_PREVIOUS_FRAME) == 0, "previous-frame SSR is not supported on peer-sharded frames");
        DFX_REQUIRE(peers->count >= 1 && peers->count <= DFX_MAX_PEERS, "peer count out of range");
        DFX_REQUIRE(c.h <= (kPeerMaxBlocks << kPeerBlockShift), "frame too tall for the peer owner table");
        DFX_REQUIRE(c.fmt == DFX_FORMAT_RGBA32F && n.fmt == DFX_FORMAT_RGBA32F, "peer-sharded planes must be RGBA32F");
        PeerArgs pa{};
        pa.count = peers->count;
        DFX_REQUIRE(peers->row_begin[0] == 0 && peers->row_begin[peers->count] == c.h, "peer strips must cover rows [0, height)");
        for (int i = 0; i <= peers->count; ++i)
        {
            const int b = peers->row_begin[i];
            DFX_REQUIRE((i == 0 || b >= peers->row_begin[i - 1]) && (b == c.h || (b & ((1 << kPeerBlockShift) - 1)) == 0),
                        "peer strip boundaries must be ascending multiples of 64");
            pa.row_begin[i] = b;
        }
        for (int i = 0; i < peers->count; ++i)
        {
            const bool empty = peers->row_begin[i + 1] == peers->row_begin[i];
            DFX_REQUIRE(empty || (peers->color[i] && peers->normal[i]), "null peer plane");
            pa.color[i]  = static_cast<const float4*>(peers->color[i]);
            pa.normal[i] = static_cast<const float4*>(peers->normal[i]);
            for (int m = 0; m < H.levels; ++m)
            {
                DFX_REQUIRE(empty || peers->hiz[m][i], "null peer Hi-Z level");
                pa.hiz[m][i] = static_cast<const float*>(peers->hiz[m][i]);
            }
        }
        Tex2 mv{nullptr, 0, 0, 0, g16 ? DFX_FORMAT_RG16F : DFX_FORMAT_RG32F};
        if (rev)
            ssr_intersect_kernel<false, true, true, false><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, c, n, r, k, bn, H, mv, orad, odir, rows.y0, rows.y1, 0, pa);
        else
            ssr_intersect_kernel<false, true, false, false><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, c, n, r, k, bn, H, mv, orad, odir, rows.y0, rows.y1, 0, pa);
    }
    else if (flags & DFX_SSR_FEATURE_FLAG_PREVIOUS_FRAME)
    {
        DFX_TEX2(mv, motion);
        DFX_SAME_SIZE(c, mv);
        DFX_REQUIRE(is16(mv) == g16, "motion must be RG16F with an RGBA16F G-buffer and RG32F with an RGBA32F one");
        if (rev)
            DFX_FMT16(g16, G16, ssr_intersect_kernel<true, false, true, G16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, c, n, r, k, bn, H, mv, orad, odir, rows.y0, rows.y1, half, NoPeerTables{}));
        else
            DFX_FMT16(g16, G16, ssr_intersect_kernel<true, false, false, G16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, c, n, r, k, bn, H, mv, orad, odir, rows.y0, rows.y1, half, NoPeerTables{}));
    }
    else
    {
        Tex2 mv{nullptr, 0, 0, 0, g16 ? DFX_FORMAT_RG16F : DFX_FORMAT_RG32F};
        if (rev)
            DFX_FMT16(g16, G16, ssr_intersect_kernel<false, false, true, G16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, c, n, r, k, bn, H, mv, orad, odir, rows.y0, rows.y1, half, NoPeerTables{}));
        else
            DFX_FMT16(g16, G16, ssr_intersect_kernel<false, false, false, G16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, c, n, r, k, bn, H, mv, orad, odir, rows.y0, rows.y1, half, NoPeerTables{}));
    }
    DFX_LAUNCHED("ssr_intersect_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssr_intersect(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs, uint32_t flags,
                                             const dfx_plane* color, const dfx_plane* normal, const dfx_plane* roughness, const dfx_plane* mask,
                                             const dfx_plane* blue_noise_xy, const dfx_pyramid* hiz, const dfx_plane* motion,
                                             const dfx_plane* out_radiance, const dfx_plane* out_raydir_pdf, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssr_intersect");
    return ssr_intersect_impl(stream, cameras_dev, attribs, flags, nullptr, color, normal, roughness, mask, blue_noise_xy, hiz, motion, out_radiance,
                              out_raydir_pdf, rows);
}

extern "C" dfx_status dfx_pass_ssr_intersect_peer(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs, uint32_t flags,
                                                  const dfx_peer_set* peers, const dfx_plane* color, const dfx_plane* normal,
                                                  const dfx_plane* roughness, const dfx_plane* mask, const dfx_plane* blue_noise_xy,
                                                  const dfx_pyramid* hiz, const dfx_plane* out_radiance, const dfx_plane* out_raydir_pdf, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssr_intersect_peer");
    DFX_REQUIRE(peers, "null argument");
    return ssr_intersect_impl(stream, cameras_dev, attribs, flags, peers, color, normal, roughness, mask, blue_noise_xy, hiz, nullptr, out_radiance,
                              out_raydir_pdf, rows);
}

extern "C" dfx_status dfx_pass_ssr_spatial(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs,
                                           const dfx_plane* roughness, const dfx_plane* mask, const dfx_plane* normal, const dfx_plane* depth,
                                           const dfx_plane* raydir_pdf, const dfx_plane* radiance, const dfx_plane* out_resolved_radiance,
                                           const dfx_plane* out_resolved_variance, const dfx_plane* out_resolved_depth, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssr_spatial");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const float, r, roughness, DFX_FORMAT_R32F);
    DFX_VIEW(const uint8_t, k, mask, DFX_FORMAT_R8U);
    DFX_TEX4(n, normal);
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    DFX_VIEW(const float4, rd, raydir_pdf, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float4, ra, radiance, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float4, orad, out_resolved_radiance, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float, ovar, out_resolved_variance, DFX_FORMAT_R32F);
    DFX_VIEW(float, odep, out_resolved_depth, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(d, r);
    DFX_SAME_SIZE(d, k);
    DFX_SAME_SIZE(d, n);
    const int half = (rd.w != d.w || rd.h != d.h) ? 1 : 0; // half-size intersect targets = FEATURE_FLAG_HALF_RESOLUTION
    DFX_REQUIRE(!half || (rd.w == d.w / 2 && rd.h == d.h / 2), "the intersect planes must have the size of the depth plane or half of it");
    DFX_SAME_SIZE(rd, ra);
    DFX_SAME_SIZE(d, orad);
    DFX_SAME_SIZE(d, ovar);
    DFX_SAME_SIZE(d, odep);
    DFX_REQUIRE(rows_ok(rows, d.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(d.w, rows);
    DFX_FMT16(is16(n), N16, ssr_spatial_kernel<N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, r, k, n, d, rd, ra, orad, ovar, odep, rows.y0, rows.y1, half));
    DFX_LAUNCHED("ssr_spatial_kernel");
    return DFX_OK;
}

static dfx_status ssr_temporal_impl(const dfx_peer_map* peers, void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs, const dfx_plane* mask,
                                            const dfx_plane* motion, const dfx_plane* hit_depth, const dfx_plane* reprojected_depth,
                                            const dfx_plane* curr_radiance, const dfx_plane* curr_variance, const dfx_plane* previous_depth,
                                            const dfx_plane* prev_radiance, const dfx_plane* prev_variance, const dfx_plane* out_radiance,
                                            const dfx_plane* out_variance, dfx_rows rows)
{
    DFX_PROFILE(stream, peers ? "ssr_temporal_peer" : "ssr_temporal");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const uint8_t, k, mask, DFX_FORMAT_R8U);
    DFX_TEX2(mv, motion);
    DFX_VIEW(const float, hd, hit_depth, DFX_FORMAT_R32F);
    DFX_VIEW(const float, cd, reprojected_depth, DFX_FORMAT_R32F);
    DFX_VIEW(const float4, cr, curr_radiance, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float, cv, curr_variance, DFX_FORMAT_R32F);
    DFX_VIEW(const float, pd, previous_depth, DFX_FORMAT_R32F);
    DFX_VIEW(const float4, pr, prev_radiance, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float, pv, prev_variance, DFX_FORMAT_R32F);
    DFX_VIEW(float4, orad, out_radiance, DFX_FORMAT_RGBA32F);
    DFX_VIEW(float, ovar, out_variance, DFX_FORMAT_R32F);
    DFX_SAME_SIZE(cd, k);
    DFX_SAME_SIZE(cd, mv);
    DFX_SAME_SIZE(cd, hd);
    DFX_SAME_SIZE(cd, cr);
    DFX_SAME_SIZE(cd, cv);
    DFX_SAME_SIZE(cd, pd);
    DFX_SAME_SIZE(cd, pr);
    DFX_SAME_SIZE(cd, pv);
    DFX_SAME_SIZE(cd, orad);
    DFX_SAME_SIZE(cd, ovar);
    DFX_REQUIRE(rows_ok(rows, cd.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(cd.w, rows);
    if (peers)
    {
        PeerMap pm;
        DFX_REQUIRE(make_peer_map(peers, cd.h, pm), "bad peer map (rank count, 64-row aligned strips, slab bases)");
        DFX_FMT16(is16(mv), M16, ssr_temporal_kernel<true, M16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, k, mv, hd, cd, cr, cv, pd, pr, pv, orad, ovar, rows.y0, rows.y1, pm));
    }
    else
        DFX_FMT16(is16(mv), M16, ssr_temporal_kernel<false, M16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, k, mv, hd, cd, cr, cv, pd, pr, pv, orad, ovar, rows.y0, rows.y1, NoPeerMap{}));
    DFX_LAUNCHED("ssr_temporal_kernel");
    return DFX_OK;
}

extern "C" dfx_status dfx_pass_ssr_temporal(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs, const dfx_plane* mask,
                                            const dfx_plane* motion, const dfx_plane* hit_depth, const dfx_plane* reprojected_depth,
                                            const dfx_plane* curr_radiance, const dfx_plane* curr_variance, const dfx_plane* previous_depth,
                                            const dfx_plane* prev_radiance, const dfx_plane* prev_variance, const dfx_plane* out_radiance,
                                            const dfx_plane* out_variance, dfx_rows rows)
{
    return ssr_temporal_impl(nullptr, stream, cameras_dev, attribs, mask, motion, hit_depth, reprojected_depth, curr_radiance, curr_variance, previous_depth, prev_radiance,
                             prev_variance, out_radiance, out_variance, rows);
}
extern "C" dfx_status dfx_pass_ssr_temporal_peer(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs, const dfx_peer_map* peers,
                                                 const dfx_plane* mask, const dfx_plane* motion, const dfx_plane* hit_depth, const dfx_plane* reprojected_depth,
                                                 const dfx_plane* curr_radiance, const dfx_plane* curr_variance, const dfx_plane* previous_depth,
                                                 const dfx_plane* prev_radiance, const dfx_plane* prev_variance, const dfx_plane* out_radiance,
                                                 const dfx_plane* out_variance, dfx_rows rows)
{
    DFX_REQUIRE(peers, "null argument");
    return ssr_temporal_impl(peers, stream, cameras_dev, attribs, mask, motion, hit_depth, reprojected_depth, curr_radiance, curr_variance, previous_depth, prev_radiance,
                             prev_variance, out_radiance, out_variance, rows);
}

extern "C" dfx_status dfx_pass_ssr_bilateral(void* stream, const dfx_camera_attribs* cameras_dev, const dfx_ssr_attribs* attribs, const dfx_plane* mask,
                                             const dfx_plane* depth, const dfx_plane* normal, const dfx_plane* roughness, const dfx_plane* radiance,
                                             const dfx_plane* variance, const dfx_plane* out, dfx_rows rows)
{
    DFX_PROFILE(stream, "ssr_bilateral");
    DFX_REQUIRE(cameras_dev && attribs, "null argument");
    DFX_VIEW(const uint8_t, k, mask, DFX_FORMAT_R8U);
    DFX_VIEW(const float, d, depth, DFX_FORMAT_R32F);
    DFX_TEX4(n, normal);
    DFX_VIEW(const float, r, roughness, DFX_FORMAT_R32F);
    DFX_VIEW(const float4, ra, radiance, DFX_FORMAT_RGBA32F);
    DFX_VIEW(const float, va, variance, DFX_FORMAT_R32F);
    DFX_VIEW(float4, o, out, DFX_FORMAT_RGBA32F);
    DFX_SAME_SIZE(d, k);
    DFX_SAME_SIZE(d, n);
    DFX_SAME_SIZE(d, r);
    DFX_SAME_SIZE(d, ra);
    DFX_SAME_SIZE(d, va);
    DFX_SAME_SIZE(d, o);
    DFX_REQUIRE(rows_ok(rows, d.h), "bad row range");
    if (rows.y1 == rows.y0) return DFX_OK;
    DFX_GRID(d.w, rows);
    DFX_FMT16(is16(n), N16, ssr_bilateral_kernel<N16><<<grid, block, 0, as_stream(stream)>>>(cameras_dev, *attribs, k, d, n, r, ra, va, o, rows.y0, rows.y1, reversed_depth(depth)));
    DFX_LAUNCHED("ssr_bilateral_kernel");
    return DFX_OK;
}
