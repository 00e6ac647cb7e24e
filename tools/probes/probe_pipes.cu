// Micro-probes of the per-SM rates the flash kernel depends on (B200, sm_100a):
//   tcgen05.ld (TMEM -> registers) and tcgen05.st bandwidth, MUFU.EX2 rate, FMA-pipe poly exp2 rate,
//   st.shared.v4 rate, and LDTM + MUFU running together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/probe_pipes tools/probes/probe_pipes.cu
// Every probe runs one CTA per SM; cycles are clock64() deltas of warp 0 around a CTA-wide barrier pair.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../petr_b200/csrc/tc_common.cuh"
using namespace petr::tc;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.f);
    const float t = x + 12582912.f;
    const float f = x - (t - 12582912.f);
    float pz = fmaf(0.05520550534129143f, f, 0.24261397123336792f);
    pz = fmaf(pz, f, 0.6932547688484192f);
    pz = fmaf(pz, f, 0.9999276995658875f);
    return __int_as_float(__float_as_int(pz) + (__float_as_int(t) << 23));
}

// mode 0: LDTM x32 loop; 1: MUFU loop; 2: both (even warps LDTM... no: every warp alternates); 3: STTM x32;
// 4: st.shared.v4; 5: poly exp2; 6: LDTM x64 (two x32 back to back before wait); 7: MUFU 3/4 + poly 1/4
__global__ void __launch_bounds__(512, 1) probe(int mode, int iters, long long* out_cycles, float* sink) {
    extern __shared__ uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col0 = ((warp >> 2) * 64) & 511;          // spread warps of a quarter over columns
    float acc = (float)lane * 1e-3f;
    uint32_t r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = lane + i;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld_x32(tb + lane_off + col0, r);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 31]);
        }
    } else if (mode == 6) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld_x32(tb + lane_off + col0, r);
            tmem_ld_x32(tb + lane_off + col0 + 32, r + 32);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 63]);
        }
    } else if (mode == 1) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = acc + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ex2f(v[i]) - 1.5f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i];
    } else if (mode == 5) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = acc + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ex2_poly(v[i]) - 1.5f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i];
    } else if (mode == 7) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = acc + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ((i & 3) == 3 ? ex2_poly(v[i]) : ex2f(v[i])) - 1.5f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i];
    } else if (mode == 2) {
        // flash-like: 64 scores from TMEM, 64 MUFU, no stores
        for (int it = 0; it < iters; ++it) {
            tmem_ld_x32(tb + lane_off + col0, r);
            tmem_ld_x32(tb + lane_off + col0 + 32, r + 32);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) acc += ex2f(__uint_as_float(r[i]));
        }
    } else if (mode == 3) {
        for (int it = 0; it < iters; ++it) {
            tmem_st_x32(tb + lane_off + col0, r);
            tmem_st_wait();
        }
    } else if (mode == 4) {
        const uint32_t base = smem_u32(smem) + (warp * 32 + lane) * 16;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                st_shared_v4(base + ((c * 16384) & 0xFFFF) + ((it & 1) << 14), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(r[5]);
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc; float* sink;
    cudaMalloc(&cyc, sms * sizeof(long long));
    cudaMalloc(&sink, sms * 1024 * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    const char* names[] = {"LDTM.x32 (4 KB per warp-instr, wait each)", "MUFU.EX2 (16 per thread-iter)",
                           "LDTM 2x x32 + 64 MUFU (flash-like)", "STTM.x32", "st.shared.v4 x8 (4 KB per warp-iter)",
                           "poly exp2 on FMA/ALU (16 per thread-iter)", "LDTM 2x x32 then wait (8 KB per warp-iter)",
                           "12 MUFU + 4 poly per thread-iter"};
    for (int mode : {0, 6, 3, 1, 5, 7, 2, 4}) {
        for (int warps : {4, 8, 12, 16}) {
            const int iters = 2000;
            probe<<<sms, warps * 32, 128 * 1024>>>(mode, iters, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
            long long h[256];
            cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
            double per_iter = avg / iters;
            double units = 0; const char* unit = "";
            if (mode == 0 || mode == 3) { units = warps * 4096.0 / per_iter; unit = "B/clk/SM"; }
            if (mode == 6) { units = warps * 8192.0 / per_iter; unit = "B/clk/SM"; }
            if (mode == 1 || mode == 5 || mode == 7) { units = warps * 32 * 16.0 / per_iter; unit = "exp2/clk/SM"; }
            if (mode == 2) { units = warps * 32 * 64.0 / per_iter; unit = "exp2/clk/SM (with 4 B of TMEM read each)"; }
            if (mode == 4) { units = warps * 4096.0 / per_iter; unit = "B/clk/SM"; }
            printf("%-48s warps %2d: %8.1f clk/iter  -> %8.2f %s\n", names[mode], warps, per_iter, units, unit);
        }
    }
    return 0;
}
