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
bool flash_exact_only();
int flash_poly();
bool gemm_staged();
bool gemm_coalesce();
bool gemm_skinny();
bool gemm_band();
void* debug_trace();
}  // namespace tuning

#define PETR_REQUIRE(cond, code, ...) \
    do { if (!(cond)) return ::petr::fail((code), __VA_ARGS__); } while (0)

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace petr
