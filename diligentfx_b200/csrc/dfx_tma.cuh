// dfx_tma.cuh — the few TMA / mbarrier primitives the tile-staging kernels use (sm_100a), behind plain functions so that the
// kernels read as ordinary code. The host-side helpers build 2-D tensor maps over pitched planes through the driver entry
// point, so the library keeps depending on libcudart only. Out-of-range parts of a box are filled with zeros.
// (First run on a B200 in round 2: profiles/r2a — a Bloom variant staging its tile through these primitives produced the same
// frames as the plain-load kernel.)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dfx
{

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // make the initialised barrier visible to the async proxy
}
// one arrival that also announces how many bytes the TMA unit will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// box of the tensor map whose first element is (x, y) (in 64-bit elements / rows; may be negative) -> dst, completing on bar
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_addr(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(x), "r"(y)
                 : "memory");
}

// Host: tensor map over a pitched plane of fp32 texels (width x height, pitch in bytes), box = box_w x box_h texels (each <= 256).
inline bool make_tensor_map_r32f(CUtensorMap* map, const void* base, int width, int height, size_t pitch_bytes, int box_w, int box_h)
{
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void*                           fn = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<encode_fn>(fn);
    }();
    if (!encode || (pitch_bytes % 16) != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0 || box_w > 256 || box_h > 256 || (box_w * 4) % 16 != 0) return false;
    const cuuint64_t dims[2]    = {cuuint64_t(width), cuuint64_t(height)};
    const cuuint64_t strides[1] = {cuuint64_t(pitch_bytes)};
    const cuuint32_t box[2]     = {cuuint32_t(box_w), cuuint32_t(box_h)};
    const cuuint32_t estr[2]    = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Host: tensor map over a pitched plane of 16-byte texels (width x height texels, pitch in bytes), box = box_w x box_h texels.
inline bool make_tensor_map_rgba32f(CUtensorMap* map, const void* base, int width, int height, size_t pitch_bytes, int box_w, int box_h)
{
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void*                            fn = nullptr;
        cudaDriverEntryPointQueryResult  st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<encode_fn>(fn);
    }();
    if (!encode || (pitch_bytes % 16) != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0 || box_w * 2 > 256 || box_h > 256) return false;
    const cuuint64_t dims[2]    = {cuuint64_t(width) * 2, cuuint64_t(height)}; // in 64-bit elements
    const cuuint64_t strides[1] = {cuuint64_t(pitch_bytes)};                   // bytes between rows
    const cuuint32_t box[2]     = {cuuint32_t(box_w) * 2, cuuint32_t(box_h)};
    const cuuint32_t estr[2]    = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace dfx
