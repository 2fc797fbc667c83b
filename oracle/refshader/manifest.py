"""REFERENCE-SHADER RUNNER — TEST INFRASTRUCTURE ONLY. The passes librefshaders.so exports: reference shader file (resolved by
base name under /root/reference/Shaders), the macros the reference's host code compiles it with, and this repository's harness.

Variant suffixes: __rev = *_OPTION_INVERTED_DEPTH (FEATURE_FLAG_REVERSED_DEPTH), __prev = SSR_OPTION_PREVIOUS_FRAME,
__half = *_OPTION_HALF_RESOLUTION. SUPPORTED_SHADER_SRV = 1 everywhere (Direct3D / Vulkan path: one mip bound per draw)."""


def _p(name, shader, harness=None, **macros):
    return {"name": name, "shader": shader, "harness": harness or (name.split("__")[0] + ".inc"), "macros": {"SUPPORTED_SHADER_SRV": 1, **macros}}


_SSR = dict(SSR_OPTION_INVERTED_DEPTH=0, SSR_OPTION_PREVIOUS_FRAME=0, SSR_OPTION_HALF_RESOLUTION=0)
_SSR_REV = dict(_SSR, SSR_OPTION_INVERTED_DEPTH=1)

PASSES = [
    # PostFXContext (PostFXContext.cpp:515)
    _p("postfx_blue_noise", "ComputeBlueNoiseTexture.fx"),
    _p("postfx_reprojected_depth", "ComputeReprojectedDepth.fx"),
    _p("postfx_closest_motion", "ComputeClosestMotion.fx", POSTFX_OPTION_INVERTED_DEPTH=0),
    _p("postfx_closest_motion__rev", "ComputeClosestMotion.fx", POSTFX_OPTION_INVERTED_DEPTH=1),
    # ScreenSpaceReflection (ScreenSpaceReflection.cpp:472-475)
    _p("ssr_hiz", "SSR_ComputeHierarchicalDepthBuffer.fx", **_SSR),
    _p("ssr_hiz__rev", "SSR_ComputeHierarchicalDepthBuffer.fx", **_SSR_REV),
    _p("ssr_mask", "SSR_ComputeStencilMaskAndExtractRoughness.fx", **_SSR),
    _p("ssr_mask__rev", "SSR_ComputeStencilMaskAndExtractRoughness.fx", **_SSR_REV),
    _p("ssr_downsample_mask", "SSR_ComputeDownsampledStencilMask.fx", **dict(_SSR, SSR_OPTION_HALF_RESOLUTION=1)),
    _p("ssr_intersect", "SSR_ComputeIntersection.fx", **_SSR),
    _p("ssr_intersect__rev", "SSR_ComputeIntersection.fx", **_SSR_REV),
    _p("ssr_intersect__prev", "SSR_ComputeIntersection.fx", **dict(_SSR, SSR_OPTION_PREVIOUS_FRAME=1)),
    _p("ssr_intersect__half", "SSR_ComputeIntersection.fx", **dict(_SSR, SSR_OPTION_HALF_RESOLUTION=1)),
    _p("ssr_spatial", "SSR_ComputeSpatialReconstruction.fx", **_SSR),
    _p("ssr_spatial__rev", "SSR_ComputeSpatialReconstruction.fx", **_SSR_REV),
    _p("ssr_spatial__half", "SSR_ComputeSpatialReconstruction.fx", **dict(_SSR, SSR_OPTION_HALF_RESOLUTION=1)),
    _p("ssr_temporal", "SSR_ComputeTemporalAccumulation.fx", **_SSR),
    _p("ssr_temporal__rev", "SSR_ComputeTemporalAccumulation.fx", **_SSR_REV),
    _p("ssr_bilateral", "SSR_ComputeBilateralCleanup.fx", **_SSR),
    _p("ssr_bilateral__rev", "SSR_ComputeBilateralCleanup.fx", **_SSR_REV),
]
