// chain_driver.cpp — drives the C++ drop-in classes the way the reference's integration does
// (Hydrogent/src/Tasks/HnPostProcessTask.cpp: Prepare :591-683, Execute :743-947): every frame
//   PostFXContext::PrepareResources -> SSAO/SSR/TAA/Bloom::PrepareResources -> PostFXContext::Execute -> SSR -> SSAO -> compose
//   -> TAA -> Bloom -> ToneMap(+sRGB).
// Host-only translation unit (g++, no CUDA headers): everything goes through include/dfx/*.hpp and the C-ABI.
//
//   chain_driver <dir> <width> <height> <frames>
// reads <dir>/f<k>_{depth,prev_depth,motion,normal,color,material}.bin (raw fp32) and f<k>_cameras.bin (2 x 576 B),
// writes <dir>/out_ldr.bin, out_ao.bin, out_ssr.bin of the LAST frame.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "dfx/GBuffer.hpp"
#include "dfx/PostProcessEffects.hpp"

using namespace Diligent;

static std::vector<char> read_file(const std::string& path)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f)
    {
        std::fprintf(stderr, "cannot open %s\n", path.c_str());
        std::exit(2);
    }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> b(static_cast<size_t>(n));
    if (std::fread(b.data(), 1, b.size(), f) != b.size()) std::exit(2);
    std::fclose(f);
    return b;
}
static void write_file(const std::string& path, const void* p, size_t n)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    std::fwrite(p, 1, n, f);
    std::fclose(f);
}

int main(int argc, char** argv)
{
    if (argc < 5) return 1;
    const std::string dir = argv[1];
    const Uint32      W = std::atoi(argv[2]), H = std::atoi(argv[3]), Frames = std::atoi(argv[4]);

    IRenderDevice  Device;
    IDeviceContext Context; // default stream

    // G-buffer planes (what the renderer would hand over)
    // ... held in a Diligent::GBuffer (Components/interface/GBuffer.hpp), like Hydrogent's HnBeginFrameTask does
    GBuffer::ElementDesc Elems[6];
    Elems[0].Format = TEX_FORMAT_D32_FLOAT, Elems[0].BindFlags = BIND_DEPTH_STENCIL | BIND_SHADER_RESOURCE, Elems[1] = Elems[0];
    Elems[2].Format = TEX_FORMAT_RG32_FLOAT, Elems[3].Format = Elems[4].Format = Elems[5].Format = TEX_FORMAT_RGBA32_FLOAT;
    GBuffer   GB{Elems, 6, &Device, W, H};
    GB.Bind(&Context, 0x3Fu, nullptr, 0x3Fu); // clears every buffer to its ClearValue (depth 1, colours 0) before the frame's data arrives
    ITexture &Depth = *GB.GetBuffer(0), &PrevDepth = *GB.GetBuffer(1), &Motion = *GB.GetBuffer(2), &Normal = *GB.GetBuffer(3), &Color = *GB.GetBuffer(4), &Material = *GB.GetBuffer(5);
    ITexture  Composed{W, H, TEX_FORMAT_RGBA32_FLOAT}, LDR{W, H, TEX_FORMAT_RGBA32_FLOAT};
    ITextureView DepthSRV{&Depth}, PrevDepthSRV{&PrevDepth}, MotionSRV{&Motion}, NormalSRV{&Normal}, ColorSRV{&Color}, MaterialSRV{&Material};
    ITextureView ComposedView{&Composed}, LDRView{&LDR};

    PostFXContext               PostFX{&Device, {}};
    ScreenSpaceReflection       SSR{&Device, {}};
    ScreenSpaceAmbientOcclusion SSAO{&Device, {}};
    TemporalAntiAliasing        TAA{&Device, {}};
    Bloom                       BloomFX{&Device, {}};
    // the classes fade their effect in with wall-clock time like the reference (AlphaInterpolation); a test wants a fixed value
    SSR.SetAlphaInterpolation(1.0f), SSAO.SetAlphaInterpolation(1.0f), BloomFX.SetAlphaInterpolation(1.0f);

    HLSL::ScreenSpaceAmbientOcclusionAttribs SSAOAttribs;
    HLSL::ScreenSpaceReflectionAttribs       SSRAttribs;
    HLSL::TemporalAntiAliasingAttribs        TAAAttribs;
    HLSL::BloomAttribs                       BloomAttribs;
    HLSL::ToneMappingAttribs                 ToneMapAttribs;

    for (Uint32 k = 0; k < Frames; ++k)
    {
        const std::string p = dir + "/f" + std::to_string(k) + "_";
        Depth.UpdateData(&Context, read_file(p + "depth.bin").data());
        PrevDepth.UpdateData(&Context, read_file(p + "prev_depth.bin").data());
        Motion.UpdateData(&Context, read_file(p + "motion.bin").data());
        Normal.UpdateData(&Context, read_file(p + "normal.bin").data());
        Color.UpdateData(&Context, read_file(p + "color.bin").data());
        Material.UpdateData(&Context, read_file(p + "material.bin").data());
        Context.WaitForIdle();
        const std::vector<char>    cams = read_file(p + "cameras.bin");
        const HLSL::CameraAttribs* pCams = reinterpret_cast<const HLSL::CameraAttribs*>(cams.data());

        // ---- Prepare
        PostFXContext::FrameDesc FrameDesc;
        FrameDesc.Index = pCams[0].uiFrameIndex, FrameDesc.Width = W, FrameDesc.Height = H, FrameDesc.OutputWidth = W, FrameDesc.OutputHeight = H;
        PostFX.PrepareResources(&Device, FrameDesc, PostFXContext::FEATURE_FLAG_NONE);
        SSAO.PrepareResources(&Device, &Context, &PostFX, ScreenSpaceAmbientOcclusion::FEATURE_FLAG_NONE);
        SSR.PrepareResources(&Device, &Context, &PostFX, ScreenSpaceReflection::FEATURE_FLAG_NONE);
        TAA.PrepareResources(&Device, &Context, &PostFX, TemporalAntiAliasing::FEATURE_FLAG_BICUBIC_FILTER);
        BloomFX.PrepareResources(&Device, &Context, &PostFX, Bloom::FEATURE_FLAG_NONE);

        // ---- Execute
        PostFXContext::RenderAttributes PostFXAttribs;
        PostFXAttribs.pDevice = &Device, PostFXAttribs.pDeviceContext = &Context;
        PostFXAttribs.pCurrDepthBufferSRV = &DepthSRV, PostFXAttribs.pPrevDepthBufferSRV = &PrevDepthSRV, PostFXAttribs.pMotionVectorsSRV = &MotionSRV;
        PostFXAttribs.pCurrCamera = &pCams[0], PostFXAttribs.pPrevCamera = &pCams[1];
        PostFX.Execute(PostFXAttribs);

        ScreenSpaceReflection::RenderAttributes SSRRenderAttribs;
        SSRRenderAttribs.pDevice = &Device, SSRRenderAttribs.pDeviceContext = &Context, SSRRenderAttribs.pPostFXContext = &PostFX;
        SSRRenderAttribs.pColorBufferSRV = &ColorSRV, SSRRenderAttribs.pDepthBufferSRV = &DepthSRV, SSRRenderAttribs.pNormalBufferSRV = &NormalSRV;
        SSRRenderAttribs.pMaterialBufferSRV = &MaterialSRV, SSRRenderAttribs.pMotionVectorsSRV = &MotionSRV, SSRRenderAttribs.pSSRAttribs = &SSRAttribs;
        SSR.Execute(SSRRenderAttribs);

        ScreenSpaceAmbientOcclusion::RenderAttributes SSAORenderAttribs;
        SSAORenderAttribs.pDevice = &Device, SSAORenderAttribs.pDeviceContext = &Context, SSAORenderAttribs.pPostFXContext = &PostFX;
        SSAORenderAttribs.pDepthBufferSRV = &DepthSRV, SSAORenderAttribs.pNormalBufferSRV = &NormalSRV, SSAORenderAttribs.pSSAOAttribs = &SSAOAttribs;
        SSAO.Execute(SSAORenderAttribs);

        ComposeSSRAndSSAO(&Context, &ColorSRV, SSR.GetSSRRadianceSRV(), SSAO.GetAmbientOcclusionSRV(), 1.0f, 1.0f, &ComposedView);

        TemporalAntiAliasing::RenderAttributes TAARenderAttribs;
        TAARenderAttribs.pDevice = &Device, TAARenderAttribs.pDeviceContext = &Context, TAARenderAttribs.pPostFXContext = &PostFX;
        TAARenderAttribs.pColorBufferSRV = &ComposedView, TAARenderAttribs.pTAAAttribs = &TAAAttribs;
        TAA.Execute(TAARenderAttribs);

        Bloom::RenderAttributes BloomRenderAttribs;
        BloomRenderAttribs.pDevice = &Device, BloomRenderAttribs.pDeviceContext = &Context, BloomRenderAttribs.pPostFXContext = &PostFX;
        BloomRenderAttribs.pColorBufferSRV = TAA.GetAccumulatedFrameSRV(), BloomRenderAttribs.pBloomAttribs = &BloomAttribs;
        BloomFX.Execute(BloomRenderAttribs);

        ToneMap(&Context, BloomFX.GetBloomTextureSRV(), &LDRView, ToneMapAttribs, 0.3f, true);
        Context.WaitForIdle();
    }

    std::vector<float> ldr(size_t(W) * H * 4), ao(size_t(W) * H), ssr(size_t(W) * H * 4);
    LDR.ReadData(&Context, ldr.data());
    SSAO.GetAmbientOcclusionSRV()->GetTexture()->ReadData(&Context, ao.data());
    SSR.GetSSRRadianceSRV()->GetTexture()->ReadData(&Context, ssr.data());
    write_file(dir + "/out_ldr.bin", ldr.data(), ldr.size() * 4);
    write_file(dir + "/out_ao.bin", ao.data(), ao.size() * 4);
    write_file(dir + "/out_ssr.bin", ssr.data(), ssr.size() * 4);
    const float j = TAA.GetJitterOffset().x;
    std::printf("chain_driver: %u frames %ux%u done, %llu kernel launches, next jitter.x %g\n", Frames, W, H, (unsigned long long)dfx_launch_count(), j);
    return 0;
}
