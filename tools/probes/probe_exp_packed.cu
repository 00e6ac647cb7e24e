// exp2 throughput per SM of the packed half-precision MUFU forms (ex2.approx.ftz.bf16x2 / .f16x2) beside the
// fp32 form, each including the conversions a softmax warp needs around them: fp32 scores in, packed bf16 P out.
//   MODE 0: 2 x ex2.f32 + cvt.rn.bf16x2.f32            (what the flash kernel does on its MUFU share)
//   MODE 1: cvt.rn.bf16x2.f32 + ex2.approx.ftz.bf16x2  (P is already packed)
//   MODE 2: cvt.rn.f16x2.f32 + ex2.approx.f16x2 (+ nothing: f16 P)
//   MODE 3: ex2.approx.ftz.bf16x2 alone on packed inputs (pure MUFU rate)
//   MODE 4: ex2.f32 alone (pure MUFU rate)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* cyc, uint32_t* sink) {
    float v[64];
    uint32_t pk[32];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = -1.f - 0.01f * (lane + i);
#pragma unroll
    for (int i = 0; i < 32; ++i) pk[i] = cvt_bf16x2(v[2 * i], v[2 * i + 1]);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int pr = 0; pr < 32; ++pr) {
            if (MODE == 0) acc ^= cvt_bf16x2(ex2(v[2 * pr]), ex2(v[2 * pr + 1]));
            if (MODE == 1) acc ^= ex2_bf16x2(cvt_bf16x2(v[2 * pr], v[2 * pr + 1]));
            if (MODE == 2) acc ^= ex2_f16x2(cvt_f16x2(v[2 * pr], v[2 * pr + 1]));
            if (MODE == 3) acc ^= ex2_bf16x2(pk[pr]);
            if (MODE == 4) acc ^= __float_as_uint(ex2(v[2 * pr])) ^ __float_as_uint(ex2(v[2 * pr + 1]));
        }
        // every input must change every iteration: ptxas hoists loop-invariant ex2 out of the loop (a first version
        // of this probe changed one input in eight and "measured" 96 exp2/clk/SM)
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = ((acc + i * 0x00010001u) & 0x007F007Fu) | 0xBF80BF80u;
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(((acc + i) & 0x007FFFFFu) | 0xBF800000u);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, int sms, long long* cyc, uint32_t* sink) {
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        probe<MODE><<<sms, warps * 32>>>(iters, cyc, sink);
        cudaDeviceSynchronize();
        long long h[256];
        cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
        printf("%-44s warps %2d: %8.1f clk/iter -> %6.2f exp2/clk/SM\n", name, warps, avg / iters, warps * 32 * 64.0 / (avg / iters));
    }
}
// accuracy of the packed form on a sweep of inputs
__global__ void accuracy(float lo, float hi, int n, float* max_rel) {
    float worst = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = lo + (hi - lo) * i / n;
        const uint32_t y = ex2_bf16x2(cvt_bf16x2(x, x));
        const float got = __uint_as_float(y << 16);
        const float want = exp2f(x);
        worst = fmaxf(worst, fabsf(got - want) / want);
    }
    atomicMax(reinterpret_cast<int*>(max_rel), __float_as_int(worst));
}
int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc; uint32_t* sink; float* rel;
    cudaMalloc(&cyc, sms * 8); cudaMalloc(&sink, sms * 512 * 4); cudaMalloc(&rel, 4);
    run<4>("ex2.f32 alone", sms, cyc, sink);
    run<3>("ex2.bf16x2 alone (packed in)", sms, cyc, sink);
    run<0>("2 x ex2.f32 + cvt.bf16x2", sms, cyc, sink);
    run<1>("cvt.bf16x2 + ex2.bf16x2", sms, cyc, sink);
    run<2>("cvt.f16x2 + ex2.f16x2", sms, cyc, sink);
    for (float lo : {-1.f, -4.f, -16.f}) {
        cudaMemset(rel, 0, 4);
        accuracy<<<1, 256>>>(lo, 0.f, 1 << 16, rel);
        float h; cudaMemcpy(&h, rel, 4, cudaMemcpyDeviceToHost);
        printf("ex2.bf16x2(bf16(x)) vs exp2f(x), x in [%g, 0]: max rel err %.4g\n", lo, h);
    }
    return 0;
}
