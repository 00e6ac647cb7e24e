// tcgen05.ld bandwidth by width (static register use only: no local-memory traffic), alone, with MUFU work in the
// same warps, and while a tcgen05.mma stream runs.  B200, one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/probe_ldtm tools/probes/probe_ldtm.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../petr_b200/csrc/tc_common.cuh"
using namespace petr::tc;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
template <int N> __device__ __forceinline__ uint32_t xor_all(const uint32_t* r) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) x ^= r[i];
    return x;
}

// mode 0: x8; 1: x16; 2: x32; 3: 2 x x32 (64 cols) one wait; 4: 64 cols + 64 MUFU; 5: x32 while MMA; 6: 64 MUFU only (static)
template <int MODE>
__global__ void __launch_bounds__(544, 1) probe(int iters, long long* out_cycles, uint32_t* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = (blockDim.x >> 5) - (MODE == 5 ? 1 : 0);
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (MODE == 5) {
        for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + i;
        fence_proxy_async_smem();
    }
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t ta = tb + ((uint32_t)((warp & 3) * 32) << 16) + (((warp >> 2) * 64) & 255);
    uint32_t acc = lane;
    __syncthreads();
    long long t0 = clock64();
    if (MODE == 5 && warp == nw) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 1, 0, 0);
            const uint64_t ad = make_smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
            const uint64_t bd = make_smem_desc(smem_u32(smem) + 16384, 16, 1024, LAYOUT_SW128);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tb + 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            if (MODE == 0) { uint32_t r[8]; tmem_ld_x8(ta, r); tmem_ld_wait(); acc += xor_all<8>(r); }
            if (MODE == 1) { uint32_t r[16]; tmem_ld_x16(ta, r); tmem_ld_wait(); acc += xor_all<16>(r); }
            if (MODE == 2 || MODE == 5) { uint32_t r[32]; tmem_ld_x32(ta, r); tmem_ld_wait(); acc += xor_all<32>(r); }
            if (MODE == 3) { uint32_t r[64]; tmem_ld_x32(ta, r); tmem_ld_x32(ta + 32, r + 32); tmem_ld_wait(); acc += xor_all<64>(r); }
            if (MODE == 4) {
                uint32_t r[64]; tmem_ld_x32(ta, r); tmem_ld_x32(ta + 32, r + 32); tmem_ld_wait();
                uint32_t x = 0;
#pragma unroll
                for (int i = 0; i < 64; ++i) x ^= __float_as_uint(ex2f(__uint_as_float(r[i])));
                acc += x;
            }
            if (MODE == 6) {
                uint32_t x = 0;
#pragma unroll
                for (int i = 0; i < 64; ++i) x ^= __float_as_uint(ex2f(__uint_as_float(acc + i)));
                acc += x;
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tcgen05_fence_before(); __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tb, 512); }
}

template <int MODE> double run(int sms, int warps, int iters, long long* cyc, uint32_t* sink) {
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe<MODE><<<sms, (warps + (MODE == 5 ? 1 : 0)) * 32, 64 * 1024>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); exit(1); }
    long long hc[256];
    cudaMemcpy(hc, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += hc[i];
    return avg / sms / iters;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, sms * sizeof(long long));
    cudaMalloc(&sink, sms * 1024 * 4);
    const int iters = 4000;
    for (int warps : {4, 8, 16}) {
        double c;
        c = run<0>(sms, warps, iters, cyc, sink); printf("warps %2d  LDTM x8  (1 KB/warp-iter): %8.1f clk/iter -> %7.1f B/clk/SM\n", warps, c, warps * 1024.0 / c);
        c = run<1>(sms, warps, iters, cyc, sink); printf("warps %2d  LDTM x16 (2 KB/warp-iter): %8.1f clk/iter -> %7.1f B/clk/SM\n", warps, c, warps * 2048.0 / c);
        c = run<2>(sms, warps, iters, cyc, sink); printf("warps %2d  LDTM x32 (4 KB/warp-iter): %8.1f clk/iter -> %7.1f B/clk/SM\n", warps, c, warps * 4096.0 / c);
        c = run<3>(sms, warps, iters, cyc, sink); printf("warps %2d  LDTM 2 x x32 (8 KB)      : %8.1f clk/iter -> %7.1f B/clk/SM\n", warps, c, warps * 8192.0 / c);
        c = run<6>(sms, warps, iters, cyc, sink); printf("warps %2d  64 MUFU per thread-iter   : %8.1f clk/iter -> %7.2f exp2/clk/SM\n", warps, c, warps * 2048.0 / c);
        c = run<4>(sms, warps, iters, cyc, sink); printf("warps %2d  LDTM 8 KB + 64 MUFU       : %8.1f clk/iter -> %7.2f exp2/clk/SM, %7.1f B/clk/SM\n", warps, c, warps * 2048.0 / c, warps * 8192.0 / c);
        c = run<5>(sms, warps, iters, cyc, sink); printf("warps %2d  LDTM x32 while MMA stream  : %8.1f clk/iter -> %7.1f B/clk/SM (MMA nominal 512 clk/iter)\n", warps, c, warps * 4096.0 / c);
    }
    return 0;
}
