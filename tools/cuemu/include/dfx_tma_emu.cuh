// cuemu — DEVELOPMENT TOOL. Host stand-in for diligentfx_b200/csrc/dfx_tma.cuh (build_emu.py puts it in its place): a tensor map
// is a plain description of the plane, a TMA tile load is a synchronous copy with zero fill outside the plane, an mbarrier is
// a phase counter that waiting threads poll while yielding to the rest of the block.
#pragma once
#include "cuda_runtime.h"

struct CUtensorMap
{
    const char* base;
    uint64_t    dim0, dim1, stride; // 64-bit elements per row, rows, bytes between rows
    uint32_t    box0, box1;
};

namespace dfx
{
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void mbar_arrive_expect_tx(uint64_t*, uint32_t) {}
inline void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while ((*bar & 1u) == parity) ::cuemu::yield();
}
inline void tma_load_2d(void* dst, const CUtensorMap* m, int x, int y, uint64_t* bar)
{
    uint64_t* d = static_cast<uint64_t*>(dst);
    for (uint32_t r = 0; r < m->box1; ++r)
        for (uint32_t c = 0; c < m->box0; ++c)
        {
            const long long gx = (long long)x + c, gy = (long long)y + r;
            uint64_t        v  = 0;
            if (gx >= 0 && gy >= 0 && (uint64_t)gx < m->dim0 && (uint64_t)gy < m->dim1) std::memcpy(&v, m->base + (size_t)gy * m->stride + (size_t)gx * 8, 8);
            d[size_t(r) * m->box0 + c] = v;
        }
    ++*bar; // phase complete
}
inline bool make_tensor_map_rgba32f(CUtensorMap* map, const void* base, int width, int height, size_t pitch_bytes, int box_w, int box_h)
{
    if ((pitch_bytes % 16) != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0 || box_w * 2 > 256 || box_h > 256) return false;
    *map = CUtensorMap{static_cast<const char*>(base), uint64_t(width) * 2, uint64_t(height), uint64_t(pitch_bytes), uint32_t(box_w) * 2, uint32_t(box_h)};
    return true;
}
} // namespace dfx
