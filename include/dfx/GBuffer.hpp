// GBuffer.hpp — Diligent::GBuffer (Components/interface/GBuffer.hpp:41-128, Components/src/GBuffer.cpp) over CUDA planes: a set of
// device planes described by ElementDesc{Format, BindFlags, ClearValue}, created / re-created by Resize(), handed to the PostProcess
// effects through GetBuffer(i). The formats are the renderer's own (colour / normal RGBA16_FLOAT, motion RG16_FLOAT, material
// RG8_UNORM, depth D32_FLOAT — Hydrogent/src/Tasks/HnBeginFrameTask.cpp:63-69): the passes of this library read them as they are.
// There is nothing to bind in a CUDA pipeline, so Bind() keeps the part of the reference's contract that has an effect here:
// clearing the buffers named by ClearMask to their ClearValue on the context's stream.
#pragma once
#include <vector>

#include "DiligentShim.hpp"

namespace Diligent
{

enum BIND_FLAGS : Uint32
{
    BIND_NONE            = 0,
    BIND_SHADER_RESOURCE = 1u << 3,
    BIND_RENDER_TARGET   = 1u << 5,
    BIND_DEPTH_STENCIL   = 1u << 6
};
DEFINE_FLAG_ENUM_OPERATORS(BIND_FLAGS)

struct OptimizedClearValue
{
    TEXTURE_FORMAT Format   = TEX_FORMAT_UNKNOWN;
    float          Color[4] = {0, 0, 0, 0};
    struct
    {
        float Depth   = 1.0f;
        Uint8 Stencil = 0;
    } DepthStencil;
};

class GBuffer
{
public:
    struct ElementDesc
    {
        TEXTURE_FORMAT      Format     = TEX_FORMAT_UNKNOWN;
        BIND_FLAGS          BindFlags  = BIND_NONE;
        OptimizedClearValue ClearValue = {};
    };

    GBuffer(const ElementDesc* Elements, size_t NumElements) : m_ElemDesc{Elements, Elements + NumElements}, m_Buffers(NumElements), m_Depth(NumElements, false)
    {
        for (size_t i = 0; i < NumElements; ++i)
        {
            DFX_DEV_CHECK_ERR(m_ElemDesc[i].Format != TEX_FORMAT_UNKNOWN, "GBuffer element format must not be TEX_FORMAT_UNKNOWN");
            // GBuffer.cpp: BIND_NONE -> depth-stencil formats get BIND_DEPTH_STENCIL | BIND_SHADER_RESOURCE, everything else BIND_RENDER_TARGET | BIND_SHADER_RESOURCE
            m_Depth[i] = (m_ElemDesc[i].BindFlags & BIND_DEPTH_STENCIL) != 0 || (m_ElemDesc[i].BindFlags == BIND_NONE && m_ElemDesc[i].ClearValue.Format == TEX_FORMAT_D32_FLOAT);
            if (m_ElemDesc[i].BindFlags == BIND_NONE) m_ElemDesc[i].BindFlags = (m_Depth[i] ? BIND_DEPTH_STENCIL : BIND_RENDER_TARGET) | BIND_SHADER_RESOURCE;
        }
    }
    GBuffer(const ElementDesc* Elements, size_t NumElements, IRenderDevice* pDevice, Uint32 Width, Uint32 Height) : GBuffer{Elements, NumElements} { Resize(pDevice, Width, Height); }

    const ElementDesc& GetElementDesc(Uint32 Index) const { return m_ElemDesc[Index]; }
    ITexture*          GetBuffer(Uint32 Index) const { return m_Buffers[Index].get(); }
    size_t             GetBufferCount() const { return m_Buffers.size(); }

    void Resize(IRenderDevice*, Uint32 Width, Uint32 Height)
    {
        if (Width == m_Width && Height == m_Height) return;
        m_Width = Width, m_Height = Height;
        for (size_t i = 0; i < m_Buffers.size(); ++i)
            m_Buffers[i] = (Width != 0 && Height != 0) ? std::make_unique<ITexture>(Width, Height, m_ElemDesc[i].Format) : nullptr;
    }

    void Bind(IDeviceContext* pContext, Uint32 BuffersMask, ITextureView* /*pDSV*/, Uint32 ClearMask = 0, const Uint32* /*RTIndices*/ = nullptr)
    {
        for (size_t i = 0; i < m_Buffers.size(); ++i)
        {
            const Uint32 bit = 1u << i;
            if (!(BuffersMask & bit) || !(ClearMask & bit) || !m_Buffers[i]) continue;
            const ElementDesc& e = m_ElemDesc[i];
            const float        d = e.ClearValue.DepthStencil.Depth;
            const float        depth[4] = {d, d, d, d};
            detail::Check(dfx_plane_fill(detail::StreamOf(pContext), &m_Buffers[i]->GetPlane(), m_Depth[i] ? depth : e.ClearValue.Color), "GBuffer::Bind (clear)");
        }
    }

private:
    std::vector<ElementDesc>               m_ElemDesc;
    std::vector<std::unique_ptr<ITexture>> m_Buffers;
    std::vector<bool>                      m_Depth;
    Uint32                                 m_Width = 0, m_Height = 0;
};

} // namespace Diligent
