// cuemu — DEVELOPMENT TOOL, not part of the product and not built by __graft_entry__.build().
//
// A host-side stand-in for the slice of CUDA that diligentfx_b200/csrc uses, so that the kernel SOURCES can be compiled
// with g++ and executed on a machine without a GPU (tools/cuemu/build_emu.py). Purpose: check an edit to a kernel against
// the oracle in seconds before spending GPU time on it. It is not a fallback: the package never loads the emulated
// library, results are not bit-identical to the GPU's (no MUFU approximations, no FMA contraction) and it is orders of
// magnitude slower. See tools/cuemu/README.md.
//
// Execution model: blocks of a grid run one after the other (optionally spread over host threads); the threads of a block
// are coroutines on their own stacks, resumed round-robin, so __syncthreads() is a yield-until-everyone-arrived.
// __shared__ is block-local static storage (thread_local per host thread), __constant__ is static storage, device memory is
// host memory, streams and events are no-ops / host timers.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- qualifiers ----
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

// ---- built-in vector types (aggregates, like CUDA's) ----
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uchar2 { unsigned char x, y; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort2 { unsigned short x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int2   make_int2(int x, int y) { return int2{x, y}; }
inline uint2  make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uchar2 make_uchar2(unsigned char x, unsigned char y) { return uchar2{x, y}; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
struct dim3
{
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- per-thread built-ins (set by the scheduler before a coroutine is resumed) ----
namespace cuemu
{
struct Builtins
{
    uint3 threadIdx, blockIdx;
    dim3  blockDim, gridDim;
};
Builtins& builtins();
void      sync_threads();
void      yield(); // give the other threads of the block a turn (used by polling waits)
void      launch(dim3 grid, dim3 block, const std::function<void()>& thread_body);
} // namespace cuemu
#define threadIdx (::cuemu::builtins().threadIdx)
#define blockIdx (::cuemu::builtins().blockIdx)
#define blockDim (::cuemu::builtins().blockDim)
#define gridDim (::cuemu::builtins().gridDim)
inline void __syncthreads() { ::cuemu::sync_threads(); }

// ---- device functions ----
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline int      min(int a, int b) { return a < b ? a : b; }
inline int      max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline float    min(float a, float b) { return std::fmin(a, b); }
inline float    max(float a, float b) { return std::fmax(a, b); }
inline float    __saturatef(float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); }
inline float    __fdividef(float a, float b) { return a / b; }
inline float    __fmul_rn(float a, float b) { return a * b; }
inline float    __fadd_rn(float a, float b) { return a + b; }
inline float    __fsub_rn(float a, float b) { return a - b; }
inline float    __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float    __frcp_rn(float a) { return 1.0f / a; }
inline float    rsqrtf(float a) { return 1.0f / std::sqrt(a); }
inline float    __cosf(float a) { return std::cos(a); }
inline float    __sinf(float a) { return std::sin(a); }
inline float    __expf(float a) { return std::exp(a); }
inline float    __logf(float a) { return std::log(a); }
inline float    __log2f(float a) { return std::log2(a); }
inline float    __exp2f(float a) { return std::exp2(a); }
inline float    __powf(float a, float b) { return std::pow(a, b); }
inline void     __sincosf(float a, float* s, float* c) { *s = std::sin(a), *c = std::cos(a); }
inline void     sincosf_(float a, float* s, float* c) { *s = std::sin(a), *c = std::cos(a); }
inline int      __float2int_rd(float a) { return int(std::floor(a)); }
inline int      __float2int_rn(float a) { return int(std::nearbyint(a)); }
inline int      __float2int_rz(float a) { return int(a); }
inline unsigned __float_as_uint(float a)
{
    unsigned u;
    std::memcpy(&u, &a, 4);
    return u;
}
inline int __float_as_int(float a)
{
    int u;
    std::memcpy(&u, &a, 4);
    return u;
}
inline float __uint_as_float(unsigned u)
{
    float a;
    std::memcpy(&a, &u, 4);
    return a;
}
inline float __int_as_float(int u)
{
    float a;
    std::memcpy(&a, &u, 4);
    return a;
}
inline int      __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __brev(unsigned v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}
// the three MUFU forms dfx_common.cuh reaches through inline PTX (build_emu.py rewrites those asm statements to these)
namespace cuemu
{
inline float ftz(float a) { return std::fpclassify(a) == FP_SUBNORMAL ? std::copysign(0.0f, a) : a; }
inline float rcp_approx_ftz(float a) { return ftz(1.0f / ftz(a)); }
inline float sqrt_approx_ftz(float a) { return ftz(std::sqrt(ftz(a))); }
inline float rsqrt_approx_ftz(float a) { return ftz(1.0f / std::sqrt(ftz(a))); }
} // namespace cuemu

// ---- runtime API: device memory is host memory, streams are synchronous ----
typedef int cudaError_t;
enum
{
    cudaSuccess                       = 0,
    cudaErrorMemoryAllocation         = 2,
    cudaErrorNotSupported             = 801,
    cudaErrorPeerAccessAlreadyEnabled = 704
};
typedef struct cuemu_stream* cudaStream_t;
struct cuemu_event
{
    std::chrono::steady_clock::time_point t;
};
typedef cuemu_event* cudaEvent_t;
enum cudaMemcpyKind
{
    cudaMemcpyHostToHost,
    cudaMemcpyHostToDevice,
    cudaMemcpyDeviceToHost,
    cudaMemcpyDeviceToDevice,
    cudaMemcpyDefault
};
struct cudaIpcMemHandle_t
{
    char reserved[64];
};
enum
{
    cudaIpcMemLazyEnablePeerAccess = 1
};
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuemu: operation not available in the emulator"; }
inline const char* cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : "cudaErrorNotSupported"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { return *d = 0, cudaSuccess; }
// CUEMU_IPC=1: allocations are POSIX shared-memory segments, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle work between the
// processes of a multi-rank run (cuemu_runtime.cpp); otherwise plain aligned heap memory (which AddressSanitizer can guard).
cudaError_t cudaMalloc(void** p, size_t bytes);
template <class T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
cudaError_t cudaFree(void* p);
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { return std::memcpy(d, s, n), cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { return std::memcpy(d, s, n), cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t row, size_t rows, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < rows; ++r) std::memcpy(static_cast<char*>(d) + r * dp, static_cast<const char*>(s) + r * sp, row);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { return std::memset(d, v, n), cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { return std::memset(d, v, n), cudaSuccess; }
inline cudaError_t cudaMemset2D(void* d, size_t pitch, int v, size_t row, size_t rows)
{
    for (size_t r = 0; r < rows; ++r) std::memset(static_cast<char*>(d) + r * pitch, v, row);
    return cudaSuccess;
}
inline cudaError_t cudaMemset2DAsync(void* d, size_t pitch, int v, size_t row, size_t rows, cudaStream_t = nullptr) { return cudaMemset2D(d, pitch, v, row, rows); }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { return *e = new cuemu_event(), cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { return delete e, cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { return e->t = std::chrono::steady_clock::now(), cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    return *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(), cudaSuccess;
}
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { return *can = std::getenv("CUEMU_IPC") ? 1 : 0, cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return std::getenv("CUEMU_IPC") ? cudaSuccess : cudaErrorNotSupported; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
