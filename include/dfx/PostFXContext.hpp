// PostFXContext.hpp — same public surface as the reference's PostProcess/Common/interface/PostFXContext.hpp:51-172,
// implemented over the C-ABI (dfx_postfx_*). See DiligentShim.hpp for what the Diligent types stand for here.
#pragma once
#include "DiligentShim.hpp"

namespace Diligent
{

class PostFXContext
{
public:
    enum FEATURE_FLAGS : Uint32
    {
        FEATURE_FLAG_NONE                 = 0u,
        FEATURE_FLAG_REVERSED_DEPTH       = 1u << 0u, // near = 1, far = 0
        FEATURE_FLAG_HALF_PRECISION_DEPTH = 1u << 1u, // accepted: a storage-format choice, planes stay fp32 here
        FEATURE_FLAG_TEMPORAL_UPSCALING   = 1u << 2u, // Bloom runs at the frame's output resolution
    };

    struct FrameDesc
    {
        Uint32 Index        = 0; // must increase by exactly 1 per frame or the effects reset their histories
        Uint32 Width        = 0;
        Uint32 Height       = 0;
        Uint32 OutputWidth  = 0;
        Uint32 OutputHeight = 0;
    };

    struct RenderAttributes
    {
        IRenderDevice*             pDevice             = nullptr;
        IRenderStateCache*         pStateCache         = nullptr; // optional, ignored
        IDeviceContext*            pDeviceContext      = nullptr;
        ITextureView*              pCurrDepthBufferSRV = nullptr; // R32_FLOAT
        ITextureView*              pPrevDepthBufferSRV = nullptr; // R32_FLOAT
        ITextureView*              pMotionVectorsSRV   = nullptr; // RG32_FLOAT
        const HLSL::CameraAttribs* pCurrCamera         = nullptr;
        const HLSL::CameraAttribs* pPrevCamera         = nullptr;
        IBuffer*                   pCameraAttribsCB    = nullptr; // must be null: the context uploads pCurrCamera / pPrevCamera itself
    };

    enum BLUE_NOISE_DIMENSION : Uint32
    {
        BLUE_NOISE_DIMENSION_XY = 0,
        BLUE_NOISE_DIMENSION_ZW,
        BLUE_NOISE_DIMENSION_COUNT
    };

    struct CreateInfo
    {
        bool EnableAsyncCreation = false; // no pipeline states to compile: accepted, has no effect
        bool PackMatrixRowMajor  = false; // matrices are always row-major, row-vector convention here
    };

    PostFXContext(IRenderDevice* pDevice, const CreateInfo& CI)
    {
        (void)pDevice, (void)CI;
        detail::Check(dfx_postfx_create(&m_Ctx), "dfx_postfx_create");
    }
    ~PostFXContext() { dfx_postfx_destroy(m_Ctx); }
    PostFXContext(const PostFXContext&)            = delete;
    PostFXContext& operator=(const PostFXContext&) = delete;

    void PrepareResources(IRenderDevice* pDevice, const FrameDesc& Desc, FEATURE_FLAGS FeatureFlags)
    {
        DFX_DEV_CHECK_ERR(pDevice != nullptr, "pDevice must not be null");
        m_FrameDesc    = Desc;
        m_FeatureFlags = FeatureFlags;
        dfx_frame_desc d{Desc.Index, Desc.Width, Desc.Height, Desc.OutputWidth, Desc.OutputHeight};
        m_Prepared = detail::Check(dfx_postfx_prepare(m_Ctx, &d, static_cast<uint32_t>(FeatureFlags)), "PostFXContext::PrepareResources");
    }

    void Execute(const RenderAttributes& RenderAttribs)
    {
        DFX_DEV_CHECK_ERR(RenderAttribs.pDevice != nullptr, "RenderAttribs.pDevice must not be null");
        DFX_DEV_CHECK_ERR(RenderAttribs.pDeviceContext != nullptr, "RenderAttribs.pDeviceContext must not be null");
        DFX_DEV_CHECK_ERR(RenderAttribs.pCurrDepthBufferSRV != nullptr, "RenderAttribs.pCurrDepthBufferSRV must not be null");
        DFX_DEV_CHECK_ERR(RenderAttribs.pPrevDepthBufferSRV != nullptr, "RenderAttribs.pPrevDepthBufferSRV must not be null");
        DFX_DEV_CHECK_ERR(RenderAttribs.pMotionVectorsSRV != nullptr, "RenderAttribs.pMotionVectorsSRV must not be null");
        DFX_DEV_CHECK_ERR(RenderAttribs.pCameraAttribsCB == nullptr, "external camera constant buffers are not supported");
        dfx_postfx_render_attribs a{};
        a.stream         = detail::StreamOf(RenderAttribs.pDeviceContext);
        a.curr_depth     = RenderAttribs.pCurrDepthBufferSRV ? RenderAttribs.pCurrDepthBufferSRV->GetPlane() : nullptr;
        a.prev_depth     = RenderAttribs.pPrevDepthBufferSRV ? RenderAttribs.pPrevDepthBufferSRV->GetPlane() : nullptr;
        a.motion_vectors = RenderAttribs.pMotionVectorsSRV ? RenderAttribs.pMotionVectorsSRV->GetPlane() : nullptr;
        a.curr_camera    = RenderAttribs.pCurrCamera;
        a.prev_camera    = RenderAttribs.pPrevCamera;
        m_PSOsReady      = detail::Check(dfx_postfx_execute(m_Ctx, &a), "PostFXContext::Execute");
    }

    bool IsPSOsReady() const { return m_PSOsReady; }

    ITextureView* Get2DBlueNoiseSRV(BLUE_NOISE_DIMENSION Dimension) const
    {
        return View(Dimension == BLUE_NOISE_DIMENSION_XY ? DFX_POSTFX_PLANE_BLUE_NOISE_XY : DFX_POSTFX_PLANE_BLUE_NOISE_ZW, m_Views[Dimension == BLUE_NOISE_DIMENSION_XY ? 0 : 1]);
    }
    ITextureView* GetReprojectedDepth() const { return View(DFX_POSTFX_PLANE_REPROJECTED_DEPTH, m_Views[2]); }
    ITextureView* GetPreviousDepth() const { return View(DFX_POSTFX_PLANE_PREVIOUS_DEPTH, m_Views[3]); }
    ITextureView* GetClosestMotionVectors() const { return View(DFX_POSTFX_PLANE_CLOSEST_MOTION, m_Views[4]); }
    IBuffer*      GetCameraAttribsCB() const
    {
        m_CameraCB = IBuffer{dfx_postfx_get_camera_attribs_dev(m_Ctx)};
        return &m_CameraCB;
    }

    FEATURE_FLAGS    GetFeatureFlags() const { return m_FeatureFlags; }
    const FrameDesc& GetFrameDesc() const { return m_FrameDesc; }

    // C-ABI handle, for the effects (they take PostFXContext* exactly like the reference's effects do).
    dfx_postfx* GetHandle() const { return m_Ctx; }

private:
    ITextureView* View(int32_t id, detail::PlaneView& v) const
    {
        dfx_plane p{};
        if (dfx_postfx_get_plane(m_Ctx, id, &p) != DFX_OK) return nullptr;
        return v.Update(p);
    }

    dfx_postfx*               m_Ctx = nullptr;
    FrameDesc                 m_FrameDesc;
    FEATURE_FLAGS             m_FeatureFlags = FEATURE_FLAG_NONE;
    bool                      m_Prepared = false, m_PSOsReady = false;
    mutable detail::PlaneView m_Views[5];
    mutable IBuffer           m_CameraCB;
};
DEFINE_FLAG_ENUM_OPERATORS(PostFXContext::FEATURE_FLAGS)

} // namespace Diligent
