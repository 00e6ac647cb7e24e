// tcgen05 GEMM: C[M,N] = epi(A[M,K] . W[N,K]^T + bias) (+ residual), bf16 operands, fp32
// accumulation in TMEM.  Lk-side work of the PETR head: K/V projection of all decoder layers,
// position_encoder / adapt_pos3d MLPs (petr_head.py:249-267, petr_transformer.py:341-360).
//
// CTA = one 128 x 128 output tile.  Warp 0: TMA producer (A and W tiles, 128B-swizzled, 3-stage
// mbarrier ring).  Warp 1: TMEM allocator + single-thread tcgen05.mma issuer.  Warps 2-5:
// epilogue (tcgen05.ld 32x32b -> registers -> bias/act/residual -> global).  ~97 KB of shared
// memory and 128 TMEM columns per CTA, so two CTAs are resident per SM and one CTA's epilogue
// overlaps the other's main loop.
#include "common.cuh"
#include "tc_common.cuh"
#include <mutex>

namespace petr {

namespace tmap {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;
static std::once_flag g_once;

int encode(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
           const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
           CUtensorMapSwizzle swizzle) {
    std::call_once(g_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<EncodeFn>(fn);
    });
    if (!g_encode) return fail(PETR_E_ARCH, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdims[5], gstrides[5];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
    CUresult r = g_encode(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PETR_E_ALIGN, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}
}  // namespace tmap

namespace {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int TILE_A_BYTES = BM * BK * 2, TILE_B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = TILE_A_BYTES + TILE_B_BYTES;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GEMM_THREADS = 192;

struct EpiParams {
    const float* bias; const float* residual; int act; void* C; int ldc; int out_mode;
    int M, N, K; int kv_Lk; int kv_B; int kv_E;
};

__global__ void __launch_bounds__(GEMM_THREADS)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    EpiParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int num_k = p.K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(accum_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, BN);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                uint8_t* a = smem + s * STAGE_BYTES;
                tma_load_2d(a, &tmA, &full_bar[s], kb * BK, m0);
                tma_load_2d(a + TILE_A_BYTES, &tmB, &full_bar[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN, 1, 0, 0);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + TILE_A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024, LAYOUT_SW128);
                    uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024, LAYOUT_SW128);
                    umma_f16_ss(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);      // frees the smem stage when these MMAs retire
            }
            umma_commit(accum_bar);              // accumulator complete
        }
    } else {
        // epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int m = m0 + row;
        mbar_wait(accum_bar, 0);
        tcgen05_fence_after();
        int kv_b = 0, kv_key = 0;
        if (p.out_mode == 2 && m < p.M) { kv_b = m / p.kv_Lk; kv_key = m - kv_b * p.kv_Lk; }
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, r);
            tmem_ld_wait();
            const int nb = n0 + c * 32;
            if (m >= p.M || nb >= p.N) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                v[j] = __uint_as_float(r[j]);
                if (p.bias && nb + j < p.N) v[j] += __ldg(p.bias + nb + j);
                if (p.act == PETR_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
            }
            const bool full = (nb + 32 <= p.N);
            if (p.residual) {
                const float* rp = p.residual + (int64_t)m * p.ldc + nb;
                if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 t = *reinterpret_cast<const float4*>(rp + j);
                        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                    }
                } else {
                    for (int j = 0; j < 32 && nb + j < p.N; ++j) v[j] += rp[j];
                }
            }
            if (p.out_mode == 0) {
                float* cp = reinterpret_cast<float*>(p.C) + (int64_t)m * p.ldc + nb;
                if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    for (int j = 0; j < 32 && nb + j < p.N; ++j) cp[j] = v[j];
                }
            } else {
                __nv_bfloat16* cp;
                bool vec;
                if (p.out_mode == 1) {
                    cp = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)m * p.ldc + nb;
                    vec = full && ((p.ldc & 7) == 0);
                } else {
                    // head-split layout [n / E][b][head][key][32]; a 32-column chunk is one head
                    const int l = nb / p.kv_E, h = (nb % p.kv_E) >> 5;
                    cp = reinterpret_cast<__nv_bfloat16*>(p.C) +
                         ((((int64_t)l * p.kv_B + kv_b) * (p.kv_E >> 5) + h) * p.kv_Lk + kv_key) * 32;
                    vec = true;
                }
                if (vec) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 t;
                        t.x = pack_bf16x2(v[j], v[j + 1]); t.y = pack_bf16x2(v[j + 2], v[j + 3]);
                        t.z = pack_bf16x2(v[j + 4], v[j + 5]); t.w = pack_bf16x2(v[j + 6], v[j + 7]);
                        *reinterpret_cast<uint4*>(cp + j) = t;
                    }
                } else {
                    for (int j = 0; j < 32 && nb + j < p.N; ++j) cp[j] = __float2bfloat16_rn(v[j]);
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

}  // namespace
}  // namespace petr

using namespace petr;

extern "C" int petr_gemm_bf16_tc(const void* A, int lda, const void* W, int ldw, const float* bias,
                                 const float* residual, int act, void* C, int ldc, int out_mode,
                                 int M, int N, int K, int kv_Lk, void* stream) {
    PETR_REQUIRE(A && W && C, PETR_E_BADARG, "petr_gemm_bf16_tc: null pointer");
    PETR_REQUIRE(M > 0 && N > 0 && K > 0, PETR_E_BADARG, "petr_gemm_bf16_tc: bad dimension");
    PETR_REQUIRE(K % BK == 0, PETR_E_SHAPE, "petr_gemm_bf16_tc: K must be a multiple of 64, got %d", K);
    PETR_REQUIRE(lda >= K && ldw >= K && (lda % 8) == 0 && (ldw % 8) == 0 && aligned16(A) && aligned16(W),
                 PETR_E_ALIGN, "petr_gemm_bf16_tc: A/W rows must be 16-byte aligned");
    PETR_REQUIRE(act == PETR_ACT_NONE || act == PETR_ACT_RELU, PETR_E_BADARG, "petr_gemm_bf16_tc: bad act");
    PETR_REQUIRE(out_mode >= 0 && out_mode <= 2, PETR_E_BADARG, "petr_gemm_bf16_tc: bad out_mode");
    EpiParams p{};
    p.bias = bias; p.residual = residual; p.act = act; p.C = C; p.ldc = ldc; p.out_mode = out_mode;
    p.M = M; p.N = N; p.K = K;
    if (out_mode == 2) {
        PETR_REQUIRE(kv_Lk > 0 && M % kv_Lk == 0 && N % 256 == 0 && !residual, PETR_E_SHAPE,
                     "petr_gemm_bf16_tc: head-split output needs M %% Lk == 0, N %% 256 == 0, no residual");
        p.kv_Lk = kv_Lk; p.kv_B = M / kv_Lk; p.kv_E = 256;
        PETR_REQUIRE(aligned16(C), PETR_E_ALIGN, "petr_gemm_bf16_tc: C must be 16-byte aligned");
    } else {
        PETR_REQUIRE(ldc >= N, PETR_E_BADARG, "petr_gemm_bf16_tc: ldc < N");
    }
    CUtensorMap tmA, tmB;
    uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)lda * 2};
    uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)ldw * 2};
    uint32_t boxA[2] = {BK, BM}, boxB[2] = {BK, BN};
    int rc = tmap::encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dA, sA, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = tmap::encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dB, sB, boxB, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        if (e != cudaSuccess) return fail((int)e, "petr_gemm_bf16_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    gemm_bf16_tc_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, (cudaStream_t)stream>>>(tmA, tmB, p);
    return after_launch("petr_gemm_bf16_tc");
}
