// REFERENCE-SHADER RUNNER — TEST INFRASTRUCTURE ONLY. Included once per translation unit, before the flattened shader text.
//
// What the reference's shaders expect from their compilation environment and do not define themselves. These names come
// from DiligentCore's HLSL prelude (Graphics/HLSL2GLSLConverterLib / GraphicsEngineD3DBase "HLSLDefinitions.fxh"), a
// dependency that is NOT vendored under /root/reference; restated here for the Direct3D conventions the oracle uses
// (SURVEY.md Appendix B): NDC z in [0, 1], texture v grows downward while NDC y grows upward.
#pragma once
#include "hlsl.hpp"

#define BOOL hlsl::bool32 // ShaderDefinitions.fxh only defines BOOL when it is not defined yet; 4 bytes, as in the C structs
#define NDC_MIN_Z 0.0f
#define F3NDC_XYZ_TO_UVD_SCALE float3(0.5f, -0.5f, 1.0f)
#define MATRIX_ELEMENT(mat, row, col) mat[row][col]
#define discard throw hlsl_pixel_discarded()

struct hlsl_pixel_discarded
{
};

// The generic call every pass of librefshaders.so exports: `int refsh_<pass>(const refsh_args*)` (0 on success).
struct refsh_plane
{
    void* data; // fp32 (or uint32 for Texture2D<uint>) interleaved, row pitch = w * ch
    int   w, h, ch;
};
struct refsh_args
{
    const refsh_plane*   in;      // shader resources, in the order the pass's harness documents (mips = consecutive planes)
    int                  n_in;
    const refsh_plane*   out;     // render targets, written in place (the caller clears or pre-fills them as the reference's host code does)
    int                  n_out;
    const void* const*   cb;      // constant buffers (the C structs of include/dfx_b200.h), in the pass's order
    const int*           cb_size; // their sizes in bytes, checked against the reference's structure definitions
    int                  n_cb;
    const int*           iparam;  // pass-specific integers (instance id, mip level, ...)
    int                  n_iparam;
    const unsigned char* mask;    // optional depth/stencil test the reference's pipeline state applies: 0 = pixel not shaded
    int                  threads;
};
