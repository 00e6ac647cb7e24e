// Micro-probes (B200, sm_100a): legacy mma.sync (HMMA) rate, tcgen05.st/ld 16x128b register<->TMEM
// mapping, tcgen05.ld shape/width variants, and tcgen05.ld bandwidth while tcgen05.mma runs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/probe_mma tools/probes/probe_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../petr_b200/csrc/tc_common.cuh"
using namespace petr::tc;

__device__ __forceinline__ void hmma16816(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void tmem_st_16x128b_x1(uint32_t taddr, uint32_t r0, uint32_t r1) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x2(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {   // 16 lanes x 64 columns: 32 regs
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16_pack(uint32_t taddr, uint32_t* r) {   // 32 columns -> 16 regs
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// ---- layout check: st.16x128b.x2 then ld.32x32b.x8 ----
__global__ void layout_kernel(uint32_t* out) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tb = tmem_slot;
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0xFFFF;
    tmem_st_x32(tb, z);                     // fill lanes 0..31 x 32 columns
    tmem_st_wait();
    uint32_t r[4] = {1000u + lane, 2000u + lane, 3000u + lane, 4000u + lane};
    tmem_st_16x128b_x2(tb, r);              // lanes 0..15, columns 0..7
    uint32_t r2[4] = {5000u + lane, 6000u + lane, 7000u + lane, 8000u + lane};
    tmem_st_16x128b_x2(tb + (16u << 16) + 8, r2);   // lanes 16..31, columns 8..15
    tmem_st_wait();
    uint32_t g[16];
    tmem_ld_x16(tb, g);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[lane * 16 + i] = g[i];
    tcgen05_fence_before(); __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tb, 32); }
}

// mode 0: HMMA; 1: LDTM 32x32b.x8; 2: LDTM 16x256b.x8 (two per 32 lanes); 3: LDTM x32 pack16; 4: LDTM x32 while MMA; 5: ldmatrix.x4
__global__ void __launch_bounds__(544, 1) probe(int mode, int iters, long long* out_cycles, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = (blockDim.x >> 5) - (mode == 4 ? 1 : 0);     // worker warps
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + i;
    fence_proxy_async_smem();
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col0 = ((warp >> 2) * 64) & 255;
    float acc = lane * 1e-3f;
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = lane + i;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 4 && warp == nw) {
        // MMA issuer: 128x256x16 bf16 MMAs back to back into columns 256..511 (never read)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 1, 0, 0);
            const uint64_t ad = make_smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
            const uint64_t bd = make_smem_desc(smem_u32(smem) + 16384, 16, 1024, LAYOUT_SW128);
            for (int it = 0; it < iters * 4; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tb + 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
        }
    } else if (mode == 0) {
        float c[8][4];
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
        uint32_t a[4] = {r[0], r[1], r[2], r[3]}, b[2] = {r[4], r[5]};
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) hmma16816(c[i], a, b);
        }
        for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][3];
    } else if (mode == 1) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld_32x32b_x8(tb + lane_off + col0 + (it & 3) * 8, r);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 7]);
        }
    } else if (mode == 2) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld_16x256b_x8(tb + lane_off + col0, r);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 31]);
        }
    } else if (mode == 3) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld_32x32b_x16_pack(tb + lane_off + col0, r);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 15]);
        }
    } else if (mode == 4) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld_x32(tb + lane_off + (col0 & 127), r);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 31]);
        }
    } else if (mode == 5) {
        const uint32_t base = smem_u32(smem) + (lane & 7) * 64 + (lane >> 3) * 16 + warp * 512;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t a0, a1, a2, a3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(base + ((j * 2048 + it * 512) & 0x7FFF)));
                acc += __uint_as_float(a0 ^ a1 ^ a2 ^ a3);
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(r[5]);
    tcgen05_fence_before(); __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc; float* sink; uint32_t* lay;
    cudaMalloc(&cyc, sms * sizeof(long long));
    cudaMalloc(&sink, sms * 1024 * sizeof(float));
    cudaMalloc(&lay, 32 * 16 * 4);
    layout_kernel<<<1, 32>>>(lay);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("layout kernel: %s\n", cudaGetErrorString(e)); return 1; }
    uint32_t h[512];
    cudaMemcpy(h, lay, sizeof(h), cudaMemcpyDeviceToHost);
    printf("TMEM after st.16x128b.x2 {1000+t,2000+t,3000+t,4000+t} at (lane 0, col 0) and {5000..8000} at (lane 16, col 8): rows = TMEM lanes, 16 columns\n");
    for (int l = 0; l < 32; ++l) { printf("lane %2d:", l); for (int c = 0; c < 16; ++c) printf(" %5u", h[l * 16 + c]); printf("\n"); }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[] = {"HMMA m16n8k16 bf16 (8 per warp-iter)", "LDTM 32x32b.x8 (1 KB)", "LDTM 16x256b.x8 (16 lanes x 64 cols, 4 KB)",
                           "LDTM 32x32b.x32.pack16 (32 cols -> 16 regs)", "LDTM 32x32b.x32 while tcgen05.mma 128x256 runs", "ldmatrix.x4 (4 per warp-iter)"};
    for (int mode : {0, 5, 1, 2, 3, 4}) {
        for (int warps : {4, 8, 16}) {
            const int iters = 2000;
            const int threads = (warps + (mode == 4 ? 1 : 0)) * 32;
            probe<<<sms, threads, 64 * 1024>>>(mode, iters, cyc, sink);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
            long long hc[256];
            cudaMemcpy(hc, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < sms; ++i) avg += hc[i]; avg /= sms;
            double per_iter = avg / iters, units = 0; const char* unit = "";
            if (mode == 0) { units = warps * 8 * 2048.0 / per_iter; unit = "MAC/clk/SM"; }
            if (mode == 1) { units = warps * 1024.0 / per_iter; unit = "B/clk/SM"; }
            if (mode == 2) { units = warps * 4096.0 / per_iter; unit = "B/clk/SM"; }
            if (mode == 3) { units = warps * 4096.0 / per_iter; unit = "TMEM B/clk/SM"; }
            if (mode == 4) { units = warps * 4096.0 / per_iter; unit = "B/clk/SM (MMA clk/iter = total; 16 MMAs of 128 clk nominal per iter)"; }
            if (mode == 5) { units = warps * 4 * 512.0 / per_iter; unit = "B/clk/SM"; }
            printf("%-52s warps %2d: %8.1f clk/iter -> %9.2f %s\n", names[mode], warps, per_iter, units, unit);
        }
    }
    return 0;
}
