// fp32 FMA throughput per SM: scalar fma.rn.f32 against packed fma.rn.f32x2 (two FMAs per instruction), 16 independent
// accumulator chains per thread.  Prints FMA lanes per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int PACKED>
__global__ void __launch_bounds__(1024, 1) probe(int iters, long long* cyc, float* sink, float a0, float b0) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    uint64_t pa[16];
    const float a = a0, b = b0;
    uint64_t ap, bp;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ap) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bp) : "f"(b));
#pragma unroll
    for (int i = 0; i < 16; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(pa[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (PACKED) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(ap), "l"(bp));
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0.f;
    if (PACKED) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(pa[i])); s += x + y; }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) s += acc[i];
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int PACKED> void run(const char* name, int sms, long long* cyc, float* sink) {
    for (int threads : {256, 512, 1024}) {
        const int iters = 4000;
        probe<PACKED><<<sms, threads>>>(iters, cyc, sink, 0.999f, 1e-3f);
        cudaDeviceSynchronize();
        long long h[256];
        cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
        printf("%-22s threads %4d: %7.1f clk/iter -> %6.1f FMA lanes/clk/SM (%d instr/thread-iter)\n", name, threads,
               avg / iters, threads * 32.0 / (avg / iters), PACKED ? 16 : 32);
    }
}
int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc; float* sink;
    cudaMalloc(&cyc, sms * 8); cudaMalloc(&sink, sms * 1024 * 4);
    run<0>("fma.rn.f32", sms, cyc, sink);
    run<1>("fma.rn.f32x2", sms, cyc, sink);
    return 0;
}
