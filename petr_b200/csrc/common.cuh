// Shared host-side helpers for the C ABI (error strings, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/petr_b200.h"

namespace petr {

char* last_error_buf();             // thread-local, 512 bytes
int64_t& launch_counter();          // thread-local

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

// Called after every kernel launch: counts it and converts launch errors to a return code.
inline int after_launch(const char* what) {
    launch_counter()++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

// process-wide tuning switches behind petr_set_tuning (capi.cu)
namespace tuning {
bool pdl();
bool flash_exact_only();
int flash_poly();
bool gemm_staged();
bool gemm_coalesce();
bool gemm_skinny();
bool gemm_band();
int gemm_debug();
void* debug_trace();
}  // namespace tuning

// packed fp32x2 arithmetic (sm_100: two fp32 FMAs per issued instruction)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float2 unpack2(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}


// Programmatic dependent launch (the query side is ~100 short kernels in one stream: at B = 1 a GEMM is one wave of 8
// tiles).  A kernel launched through launch_pdl may be scheduled while its predecessor drains; pdl_wait() -- placed before
// the first access to anything a predecessor wrote or still reads -- blocks until the predecessor has completed and its
// memory is visible; pdl_launch_dependents() lets the successor's CTAs be scheduled as soon as resources free up.
// Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tuning::pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define PETR_REQUIRE(cond, code, ...) \
    do { if (!(cond)) return ::petr::fail((code), __VA_ARGS__); } while (0)

// Counter-based dropout decision shared by forward and backward kernels: a 32-bit integer hash (two multiply /
// xor-shift rounds, the "lowbias32" constants) of the element index and the seed, both folded to 32 bits first, compared
// as a 24-bit integer against p.  ~12 integer instructions per element: the attention kernels evaluate it for every
// (query, key) pair, where the 64-bit splitmix it replaces cost a quarter of their issue slots.
__host__ __device__ inline bool dropout_keep(uint64_t seed, uint64_t idx, float p) {
    const uint32_t s32 = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0xC2B2AE35u);         // loop-invariant
    uint32_t x = (uint32_t)idx * 0x9E3779B1u + (uint32_t)(idx >> 32) * 0x85EBCA77u + s32;
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return (x >> 8) >= (uint32_t)(p * 16777216.f);
}

// Ampere-style asynchronous copies: the next key / query tile streams into the other shared-memory buffer while the
// FMAs of the current one run (bytes beyond `valid` are zero-filled, nothing is read when valid is false).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool valid) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace petr
