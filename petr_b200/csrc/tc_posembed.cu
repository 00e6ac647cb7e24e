// Fused 3D position embedding (PETRHead.position_embeding, petr_head.py:286-334) on tcgen05:
//
//   frustum (u*d, v*d, d, 1)  ->  img2lidar 4x4  ->  position_range normalise  ->  inverse_sigmoid
//   ->  Conv1x1(192 -> 1024) + ReLU  ->  Conv1x1(1024 -> 256)  (+ bias, + optional fp32 residual)
//
// in ONE kernel per 128-token tile: the (B,N,W,H,D,4) coordinate tensor, the 192-channel MLP input
// and the 1024-wide hidden activation never touch HBM.  Per CTA:
//   * warps 2-5 (one thread per token) evaluate the 64 depth bins of their pixel, write the 192 bf16
//     features straight into shared memory in the UMMA K-major 128B-swizzled layout (A operand);
//   * GEMM1 runs in 8 chunks of 128 hidden units into a double-buffered TMEM accumulator; the same
//     warps apply bias + ReLU, convert to bf16 and store the chunk to shared memory as the A operand
//     of GEMM2, which accumulates the 128 x 256 output in TMEM across the 8 chunks;
//   * warp 0 streams the W1 / W2 blocks with TMA through a 7-stage 16 KB ring (weights stay L2
//     resident; 896 KB are re-streamed per tile, which is what bounds this kernel), warp 1 issues
//     tcgen05.mma (all 128 x 128 x 16).
// The bf16 path rounds the features to bf16 anyway, so the coordinate arithmetic uses fast fp32
// (FMA contraction, __fdividef, __logf); the bit-exact coordinates live in petr_frustum_features.
#include "common.cuh"
#include "tc_common.cuh"
#include <math.h>


namespace petr {
namespace {

using namespace tc;

constexpr int TM = 128;                 // tokens per CTA
constexpr int K1 = 192;                 // 3 * depth_num
constexpr int HID = 1024, EOUT = 256;
constexpr int HC = 128;                 // hidden units per chunk
constexpr int NCH = HID / HC;           // 8
constexpr int A1_BYTES = TM * K1 * 2;                   // 48 KB: 3 K-atoms of 16 KB
constexpr int WST = 7;                                  // weight ring: 7 stages of [128 rows x 64 k] = 16 KB
constexpr int WST_BYTES = 128 * 64 * 2;
constexpr int HS_BYTES = TM * HC * 2;                   // 32 KB: 2 K-atoms of 16 KB
constexpr int OFF_A1 = 0;
constexpr int OFF_WR = OFF_A1 + A1_BYTES;               // weight ring (W1 k-blocks and W2 half k-blocks, in MMA order)
constexpr int OFF_HS = OFF_WR + WST * WST_BYTES;
constexpr int OFF_B1 = OFF_HS + HS_BYTES;               // 1024 floats
constexpr int OFF_B2 = OFF_B1 + HID * 4;                // 256 floats
constexpr int OFF_BAR = OFF_B2 + EOUT * 4;
constexpr int PE_SMEM = OFF_BAR + 256 + 1024;
constexpr int PE_THREADS = 192;
constexpr int TM_H = 0;                 // H chunk accumulators at columns [0,128) and [128,256)
constexpr int TM_O = 256;               // output accumulator, 256 columns

struct PeParams {
    const float* img2lidar;             // [B*N, 16]
    int B, N, H, W, pad_h, pad_w, lid;
    float depth_start, bin_size;
    float scale[3], shift[3];           // x_norm = X * scale + shift  (= (X - lo) / span)
    const float* b1; const float* b2; const float* residual; float* out;
    const float* residual2; int res2_rows; __nv_bfloat16* out_bf16;   // + residual2[t % res2_rows]; optional bf16 store
    int64_t T;
};


__device__ __forceinline__ float inv_sigmoid_fast(float x) {
    x = fminf(fmaxf(x, 0.f), 1.f);
    const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.f - x, 1e-5f);
    return __logf(__fdividef(x1, x2));
}

__global__ void __launch_bounds__(PE_THREADS, 1)
position_mlp_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                          PeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a1_ready = bars;            // 128 producers
    uint64_t* w_full = bars + 1;          // [WST]
    uint64_t* w_empty = w_full + WST;     // [WST]
    uint64_t* h_full = w_empty + WST;     // [2]
    uint64_t* h_free = h_full + 2;        // [2], 128
    uint64_t* hs_ready = h_free + 2;      // 128
    uint64_t* hs_free = hs_ready + 1;
    uint64_t* o_full = hs_free + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
    float* sb1 = reinterpret_cast<float*>(smem + OFF_B1);
    float* sb2 = reinterpret_cast<float*>(smem + OFF_B2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas
    const int64_t tile0 = (int64_t)blockIdx.x * TM;

    for (int i = threadIdx.x; i < HID; i += PE_THREADS) sb1[i] = p.b1[i];
    for (int i = threadIdx.x; i < EOUT; i += PE_THREADS) sb2[i] = p.b2[i];
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(a1_ready, 128);
            for (int i = 0; i < WST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&h_full[i], 1); mbar_init(&h_free[i], 128); }
            mbar_init(hs_ready, 128); mbar_init(hs_free, 1); mbar_init(o_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA: stream the weights through a 7-stage ring, in the MMA's order ----------------
        // per hidden chunk c: 3 k-blocks of W1[c] (GEMM1), then the 4 half k-blocks of W2[:, c-1] (GEMM2)
        {                                          // whole warp walks the loop, the elected lane issues
            const bool leader = elect_one();
            uint32_t it = 0;
            auto load = [&](const CUtensorMap* m, int c0, int c1) {
                const int st = it % WST;
                mbar_wait(&w_empty[st], ((it / WST) & 1) ^ 1);
                if (leader) {
                    mbar_arrive_expect_tx(&w_full[st], WST_BYTES);
                    tma_load_2d(smem + OFF_WR + st * WST_BYTES, m, &w_full[st], c0, c1);
                }
                __syncwarp();
                ++it;
            };
            auto load_w2 = [&](int c) {
                for (int kb = 0; kb < HC / 64; ++kb)
                    for (int nh = 0; nh < 2; ++nh) load(&tmW2, c * HC + kb * 64, nh * 128);
            };
            for (int c = 0; c < NCH; ++c) {
                for (int kb = 0; kb < K1 / 64; ++kb) load(&tmW1, kb * 64, c * HC);
                if (c > 0) load_w2(c - 1);
            }
            load_w2(NCH - 1);
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        {                                          // whole warp; tcgen05.mma / commit by the elected lane
            const bool leader = elect_one();
            constexpr uint32_t idesc = make_idesc(TM, 128, 1, 0, 0);       // every MMA here is 128 x 128 x 16
            const uint64_t a1d = make_smem_desc(smem_u32(smem + OFF_A1), 16, 1024, LAYOUT_SW128);
            const uint64_t wrd = make_smem_desc(smem_u32(smem + OFF_WR), 16, 1024, LAYOUT_SW128);
            const uint64_t hsd = make_smem_desc(smem_u32(smem + OFF_HS), 16, 1024, LAYOUT_SW128);
            uint32_t it = 0;
            auto next_stage = [&]() -> uint64_t {
                const int st = it % WST;
                mbar_wait(&w_full[st], (it / WST) & 1);
                tcgen05_fence_after();
                ++it;
                return wrd + (uint64_t)((st * WST_BYTES) >> 4);
            };
            auto release = [&]() { if (leader) umma_commit(&w_empty[(it - 1) % WST]); __syncwarp(); };
            auto gemm2 = [&](int c) {            // O[128 x 256] += Hs[128 x 128] W2[:, c]^T, two 128-column halves
                mbar_wait(hs_ready, c & 1);
                tcgen05_fence_after();
                for (int kb = 0; kb < HC / 64; ++kb)
                    for (int nh = 0; nh < 2; ++nh) {
                        const uint64_t bd = next_stage();
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16_ss(tmem_base + TM_O + nh * 128, hsd + (uint64_t)((kb * 16384 + k * 32) >> 4),
                                            bd + (uint64_t)((k * 32) >> 4), idesc, (c | kb | k) ? 1u : 0u);
                        }
                        release();
                    }
                if (leader) umma_commit(hs_free);
                __syncwarp();
            };
            mbar_wait(a1_ready, 0);
            for (int c = 0; c < NCH; ++c) {
                mbar_wait(&h_free[c & 1], ((c >> 1) & 1) ^ 1);
                for (int kb = 0; kb < K1 / 64; ++kb) {
                    const uint64_t bd = next_stage();
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ss(tmem_base + TM_H + (c & 1) * HC, a1d + (uint64_t)((kb * 16384 + k * 32) >> 4),
                                        bd + (uint64_t)((k * 32) >> 4), idesc, (kb | k) ? 1u : 0u);
                    }
                    release();
                }
                if (leader) umma_commit(&h_full[c & 1]);
                __syncwarp();
                if (c > 0) gemm2(c - 1);
            }
            gemm2(NCH - 1);
            if (leader) umma_commit(o_full);
            __syncwarp();
        }
    } else {
        // ---------------- producer (features) + epilogues: thread <-> token ----------------
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const int64_t t = tile0 + row;
        uint8_t* arow = smem + OFF_A1 + (row >> 3) * 1024 + (row & 7) * 128;
        {
            const bool live = t < p.T;
            float M[12], u = 0.f, v = 0.f;
#pragma unroll
            for (int i = 0; i < 12; ++i) M[i] = 0.f;
            if (live) {
                const int w = (int)(t % p.W);
                const int h = (int)((t / p.W) % p.H);
                const int64_t bn = t / ((int64_t)p.W * p.H);
                const float4* mp = reinterpret_cast<const float4*>(p.img2lidar + bn * 16);
                const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
                M[0] = m0.x; M[1] = m0.y; M[2] = m0.z; M[3] = m0.w;
                M[4] = m1.x; M[5] = m1.y; M[6] = m1.z; M[7] = m1.w;
                M[8] = m2.x; M[9] = m2.y; M[10] = m2.z; M[11] = m2.w;
                u = (float)w * (float)p.pad_w / (float)p.W;
                v = (float)h * (float)p.pad_h / (float)p.H;
            }
#pragma unroll 1
            for (int g = 0; g < 8; ++g) {                  // 8 depth bins -> 24 channels -> 3 x 16-byte chunks
                float f[24];
#pragma unroll
                for (int dd = 0; dd < 8; ++dd) {
                    const float fd = (float)(g * 8 + dd);
                    const float depth = p.lid ? p.depth_start + p.bin_size * fd * (fd + 1.f)
                                              : p.depth_start + p.bin_size * fd;
                    const float dm = fmaxf(depth, 1e-5f);
                    const float p0 = u * dm, p1 = v * dm;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float X = M[4 * i] * p0 + M[4 * i + 1] * p1 + M[4 * i + 2] * depth + M[4 * i + 3];
                        // rows past T run on a zero matrix (finite values) and are never stored: no branch
                        // here, so the 24 independent chains of a group interleave (MUFU latency hidden)
                        f[dd * 3 + i] = inv_sigmoid_fast(X * p.scale[i] + p.shift[i]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int gc = 3 * g + j;                 // 16-byte chunk index along the 192 channels
                    uint8_t* dst = arow + (gc >> 3) * 16384 + (((gc & 7) ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = make_uint4(
                        pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                        pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(a1_ready);
        }
        // ---- per-chunk epilogue of GEMM1: bias + ReLU -> bf16 -> shared-memory A operand of GEMM2 ----
        uint8_t* hrow = smem + OFF_HS + (row >> 3) * 1024 + (row & 7) * 128;
        for (int c = 0; c < NCH; ++c) {
            mbar_wait(&h_full[c & 1], (c >> 1) & 1);
            tcgen05_fence_after();
            uint32_t r[128];
            const uint32_t ha = tmem_base + lane_off + TM_H + (c & 1) * HC;
            tmem_ld_x32(ha, r); tmem_ld_x32(ha + 32, r + 32); tmem_ld_x32(ha + 64, r + 64); tmem_ld_x32(ha + 96, r + 96);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&h_free[c & 1]);
            uint32_t pk[64];
            const float* bb = sb1 + c * HC;
#pragma unroll
            for (int i = 0; i < 128; i += 2)
                pk[i >> 1] = pack_bf16x2(fmaxf(__uint_as_float(r[i]) + bb[i], 0.f),
                                         fmaxf(__uint_as_float(r[i + 1]) + bb[i + 1], 0.f));
            mbar_wait(hs_free, (c & 1) ^ 1);               // GEMM2(c-1) has consumed the previous chunk
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                uint8_t* dst = hrow + (q >> 3) * 16384 + (((q & 7) ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(hs_ready);
        }
        // ---- final epilogue: O + b2 (+ residual) -> fp32 tokens, staged through shared memory ----
        // Every MMA has retired, so the A1 / weight-ring area (160 KB) is free: each warp parks its
        // 32 rows x 256 fp32 there (XOR-swizzled 16-byte slots: conflict-free both ways) and then
        // streams them out row by row, so the residual loads and the stores are fully coalesced
        // 512-byte segments instead of 32 scattered 16-byte pieces per instruction.
        mbar_wait(o_full, 0);
        tcgen05_fence_after();
        float4* stage = reinterpret_cast<float4*>(smem + OFF_A1) + (size_t)quarter * 32 * 64;     // [32 rows][64 float4]
#pragma unroll 1
        for (int c = 0; c < EOUT / 32; ++c) {
            uint32_t r[32];
            tmem_ld_x32(tmem_base + lane_off + TM_O + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int slot = (c * 8 + j) ^ lane;
                stage[lane * 64 + slot] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                      __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
        }
        __syncwarp();
        const int64_t t0w = tile0 + quarter * 32;
        int r2 = p.residual2 ? (int)(t0w % p.res2_rows) : 0;       // row of the broadcast residual (one 64-bit mod)
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr, r2 = (r2 + 1 == p.res2_rows) ? 0 : r2 + 1) {
            const int64_t tt = t0w + rr;
            if (tt >= p.T) break;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int j = lane + 32 * i;                       // float4 index along the 256 outputs
                float4 o = stage[rr * 64 + (j ^ rr)];
                const float4 bz = *reinterpret_cast<const float4*>(sb2 + 4 * j);
                o.x += bz.x; o.y += bz.y; o.z += bz.z; o.w += bz.w;
                if (p.residual) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4*>(p.residual + tt * EOUT) + j);
                    o.x += q4.x; o.y += q4.y; o.z += q4.z; o.w += q4.w;
                }
                if (p.residual2) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4*>(p.residual2 + (int64_t)r2 * EOUT) + j);
                    o.x += q4.x; o.y += q4.y; o.z += q4.z; o.w += q4.w;
                }
                if (p.out) reinterpret_cast<float4*>(p.out + tt * EOUT)[j] = o;
                if (p.out_bf16)
                    reinterpret_cast<uint2*>(p.out_bf16 + tt * EOUT)[j] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace
}  // namespace petr

using namespace petr;

extern "C" int petr_position_mlp_fused(const float* img2lidar, int B, int N, int H, int W, int D,
                                       int pad_h, int pad_w, int lid, float depth_start, float bin_size,
                                       const float* h_range_lo, const float* h_range_span,
                                       const void* W1, const float* b1, const void* W2, const float* b2,
                                       int hidden, int embed, const float* residual, const float* residual2,
                                       int res2_rows, float* out, void* out_bf16, void* stream) {
    PETR_REQUIRE(img2lidar && W1 && b1 && W2 && b2 && (out || out_bf16) && h_range_lo && h_range_span, PETR_E_BADARG,
                 "petr_position_mlp_fused: null pointer");
    if (residual2) PETR_REQUIRE(res2_rows > 0 && aligned16(residual2), PETR_E_BADARG, "petr_position_mlp_fused: residual2");
    PETR_REQUIRE(B > 0 && N > 0 && H > 0 && W > 0, PETR_E_BADARG, "petr_position_mlp_fused: bad dimension");
    PETR_REQUIRE(D == 64 && hidden == HID && embed == EOUT, PETR_E_SHAPE,
                 "petr_position_mlp_fused: specialised for depth_num 64, hidden 1024, embed 256 (got %d, %d, %d)",
                 D, hidden, embed);
    PETR_REQUIRE(aligned16(W1) && aligned16(W2) && (!out || aligned16(out)) && (!out_bf16 || aligned16(out_bf16)) &&
                 aligned16(img2lidar) && (!residual || aligned16(residual)), PETR_E_ALIGN,
                 "petr_position_mlp_fused: 16-byte alignment");
    PeParams p{};
    p.img2lidar = img2lidar; p.B = B; p.N = N; p.H = H; p.W = W; p.pad_h = pad_h; p.pad_w = pad_w; p.lid = lid;
    p.depth_start = depth_start; p.bin_size = bin_size;
    for (int i = 0; i < 3; ++i) { p.scale[i] = 1.f / h_range_span[i]; p.shift[i] = -h_range_lo[i] / h_range_span[i]; }
    p.b1 = b1; p.b2 = b2; p.residual = residual; p.out = out;
    p.residual2 = residual2; p.res2_rows = res2_rows; p.out_bf16 = (__nv_bfloat16*)out_bf16;
    p.T = (int64_t)B * N * H * W;
    CUtensorMap tmW1, tmW2;
    {
        uint64_t d[2] = {K1, HID}, s[1] = {K1 * 2};
        uint32_t box[2] = {64, HC};
        int rc = tmap::encode(&tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W1, d, s, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        uint64_t d[2] = {HID, EOUT}, s[1] = {HID * 2};
        uint32_t box[2] = {64, 128};
        int rc = tmap::encode(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W2, d, s, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute(position_mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM);
    if (e != cudaSuccess) return fail((int)e, "petr_position_mlp_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const unsigned grid = (unsigned)((p.T + TM - 1) / TM);
    position_mlp_fused_kernel<<<grid, PE_THREADS, PE_SMEM, (cudaStream_t)stream>>>(tmW1, tmW2, p);
    return after_launch("petr_position_mlp_fused");
}
