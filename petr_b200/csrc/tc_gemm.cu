// tcgen05 GEMM: C[M,N] = alpha * epi(A[M,K] . W[N,K]^T + bias) (+ residual), bf16 operands, fp32
// accumulation in TMEM.  Used for the Lk-side work of the PETR head (K/V projection of all decoder
// layers, position_encoder / adapt_pos3d MLPs; petr_head.py:249-267, petr_transformer.py:341-360)
// and, through the bf16x3 split (A = [hi|lo|hi], W = [hi|hi|lo] concatenated along K), for the
// fp32-class query-side projections / FFN.
//
// Persistent kernel, one CTA per SM, 128 x 256 output tiles (n fastest so an A tile is reused from
// L2 by neighbouring CTAs).  Warp 0: TMA producer (128B-swizzled tiles, 4-stage mbarrier ring).
// Warp 1: TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x256x16).  Warps 2-5:
// epilogue (tcgen05.ld 32x32b -> registers -> bias/act/residual -> global).  The fp32 accumulator
// is double buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the main loop
// of tile i+1.
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

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int TILE_A_BYTES = BM * BK * 2, TILE_B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = TILE_A_BYTES + TILE_B_BYTES;                 // 48 KB
constexpr int STAGE_OUT_BYTES = 8 * 2048;      // one 32-row x 64-byte staging slab per epilogue warp (out_mode 2)
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*bias tiles*/ + STAGE_OUT_BYTES;
constexpr int GEMM_THREADS = 320;      // TMA warp, MMA warp, 8 epilogue warps

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct EpiParams {
    const float* bias; const float* residual; const float* gate; int act; float alpha;
    void* C; void* C2; int ldc; int out_mode;
    int M, N, K; int a_k_wrap; int kv_Lk; int kv_B; int kv_E;
    int tiles_m, tiles_n, band;
    long long* trace;  // diagnostic clock64 stamps of CTA 0 (petr_debug_set_trace), normally null
    int coalesce;      // transposing (coalesced) epilogue allowed: 16-byte aligned bases, ldc % 4 == 0, N % 4 == 0
    int ksplit;        // out_mode 0: split-K parts (part s writes C + s * M * ldc; bias / residual in part 0)
    int tma_out;       // out_modes 3 / 4: split halves leave through swizzled slabs + TMA tensor stores (tmC / tmD)
    int dbg;           // diagnostic (petr_set_tuning gemm_debug): 1 = out_mode 2 skips its global stores
    int staged;        // out_mode 2: stage each warp's 32 x 32 bf16 chunk in smem and write it with one bulk copy
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* cp, const float* v) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        uint4 t;
        t.x = pack_bf16x2(v[j], v[j + 1]); t.y = pack_bf16x2(v[j + 2], v[j + 3]);
        t.z = pack_bf16x2(v[j + 4], v[j + 5]); t.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(cp + j) = t;
    }
}

// 32 fp32 values -> bf16 hi row and bf16 residual row (bf16x3 operand pair), 4 x 16-byte stores each
__device__ __forceinline__ void store_split_bf16x32(__nv_bfloat16* hp, __nv_bfloat16* lp, const float* v) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        uint4 h, l;
        split_bf16x2(v[j], v[j + 1], h.x, l.x); split_bf16x2(v[j + 2], v[j + 3], h.y, l.y);
        split_bf16x2(v[j + 4], v[j + 5], h.z, l.z); split_bf16x2(v[j + 6], v[j + 7], h.w, l.w);
        *reinterpret_cast<uint4*>(hp + j) = h;
        *reinterpret_cast<uint4*>(lp + j) = l;
    }
}

// tile index -> (m tile, n tile).  band == 0: n fastest (an A tile is reused from L2 by neighbouring CTAs).  band > 0
// (store-bound GEMMs with several n tiles, i.e. the K / V projection): bands of `band` m tiles, inside a band m fastest,
// so one wave of CTAs writes CONSECUTIVE key tiles of one layer (contiguous DRAM pages) and the band's A tiles stay in
// L2 for its tiles_n passes.
struct TileMap {
    int tiles_m, tiles_n, band;
    __device__ __forceinline__ void operator()(int tile, int& mt, int& nt) const {
        if (band <= 0) { mt = tile / tiles_n; nt = tile - mt * tiles_n; return; }
        const int per_band = band * tiles_n;
        const int bd = tile / per_band, r = tile - bd * per_band;
        const int rows = min(band, tiles_m - bd * band);           // the last band may be short
        nt = r / rows;
        mt = bd * band + (r - nt * rows);
    }
};

#define GEMM_TR(role, t, ev) do { if (p.trace && blockIdx.x == 0 && (t) < 64) p.trace[((role) * 64 + (t)) * 8 + (ev)] = clock64(); } while (0)

// One instantiation per output mode: with all five store paths in one kernel the epilogue alone was ~70 KB of
// SASS, far beyond the 32 KB instruction cache (measured: 3 300 clk per 32-column chunk in single-tile GEMMs).
// TF32 = true: fp32 operands through kind::tf32 (32 elements = 128 bytes per K block: the same byte geometry, the same
// four MMAs per stage).  With A = [hi | lo] and W = [W_hi | W_hi | W_lo] (hi = value rounded to tf32, lo = remainder)
// this is the 3xTF32 product hi*hi + lo*hi + hi*lo, accurate to ~2^-21: the fp32-parity path on the tensor cores.
// out_mode 6 (TF32 only): fp32 [hi | lo] split rows of the result (the A operand of the next 3xTF32 GEMM).
template <int OM, bool TF32 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, EpiParams p) {
    constexpr int BKE = TF32 ? 32 : BK;                    // K elements per 128-byte block
    // out_mode 2 (head-split K / V projection, K = 256: four k-blocks per tile, epilogue-bound -- measured: TMEM loads
    // only 0.20 ms, + bias / pack 0.31, + staging through shared memory 0.45, + the global stores 0.48): a 3-stage
    // operand ring, the fourth stage's 48 KB + the old 16 KB = four 2 KB store slabs per epilogue warp (one per
    // 32-column chunk of the tile), written in the 64-byte TMA swizzle (conflict-free: the linear layout had a 4-way
    // bank conflict on every st.shared.v4) and sent with cp.async.bulk.tensor; ONE proxy fence per two chunks.
    // out_modes 3 (bf16 [hi | lo] split rows: FFN1, 68 us against a 34 us MMA floor at B = 16) and 4 (head-split hi / lo
    // pair: the self-attention projections) take the same route when the host set p.tma_out: two slabs (hi, lo) per
    // chunk, four per pair of chunks, reused every pair after a wait on the previous pair's stores.
    constexpr int NST = (OM == 2 || OM == 3 || OM == 4) ? 3 : STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;      // 2
    uint64_t* acc_empty = acc_full + 2;           // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* sbias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);      // [2][BN]
    uint8_t* stage_out = smem + STAGES * STAGE_BYTES + 256 + 2048;                   // [8 warps][32 rows][64 B]

    // warp index through a shuffle: ptxas then treats it as warp-uniform and keeps the producer / MMA issuer
    // loops (descriptors, barrier addresses) in uniform registers instead of R2UR / ELECT loops per instruction
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int num_k = p.K / BKE;
    const int ksplit = (OM == 0) ? p.ksplit : 1;
    const int num_tiles = p.tiles_m * p.tiles_n * ksplit;        // tile index = (output tile) * ksplit + K part
    const TileMap tmap_{p.tiles_m, p.tiles_n, p.band};

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 256); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // TMA producer: the whole warp walks the loop (waits by all lanes), one elected lane issues
        const bool leader = elect_one();
        uint32_t it = 0;                                   // global k-block counter of this CTA
        uint32_t ttile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ttile) {
            int mt, nt;
            const int ks = tile % ksplit;
            tmap_(tile / ksplit, mt, nt);
            const int m0 = mt * BM, n0 = nt * BN;
            const int kb_begin = (int)((int64_t)num_k * ks / ksplit), kb_end = (int)((int64_t)num_k * (ks + 1) / ksplit);
            for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (leader) {
                    if (kb - kb_begin < 4) GEMM_TR(0, ttile, kb - kb_begin);
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    uint8_t* a = smem + s * STAGE_BYTES;
                    int ka = kb * BKE;
                    if (ka >= p.a_k_wrap) ka -= p.a_k_wrap;      // A = [hi | lo | hi] without a third copy
                    tma_load_2d(a, &tmA, &full_bar[s], ka, m0);
                    tma_load_2d(a + TILE_A_BYTES, &tmB, &full_bar[s], kb * BKE, n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(BM, BN, TF32 ? 2 : 1, 0, 0);
        const uint64_t desc0 = make_smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t buf = tcount & 1;
            mbar_wait(&acc_empty[buf], ((tcount >> 1) & 1) ^ 1);   // epilogue drained this accumulator
            tcgen05_fence_after();
            if (leader) GEMM_TR(1, tcount, 0);
            const uint32_t d = tmem_base + buf * BN;
            const int ks = tile % ksplit;
            const int kb_begin = (int)((int64_t)num_k * ks / ksplit), kb_end = (int)((int64_t)num_k * (ks + 1) / ksplit);
            for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                const int s = it % NST;
                mbar_wait(&full_bar[s], (it / NST) & 1);
                tcgen05_fence_after();
                if (leader) {
                    if (kb - kb_begin < 4) GEMM_TR(1, tcount, 1 + kb - kb_begin);
                    const uint64_t ad = desc0 + (uint64_t)((s * STAGE_BYTES) >> 4);
                    const uint64_t bd = ad + (uint64_t)(TILE_A_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {          // 32 bytes of K per MMA: 16 bf16 or 8 tf32
                        if constexpr (TF32) umma_tf32_ss(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, ((kb - kb_begin) | k) ? 1u : 0u);
                        else umma_f16_ss(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, ((kb - kb_begin) | k) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);                  // smem stage free when these retire
                }
                __syncwarp();
            }
            if (leader) { umma_commit(&acc_full[buf]); GEMM_TR(1, tcount, 5); }
            __syncwarp();
        }
    } else {
        // epilogue: 8 warps; warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) and takes column
        // half (w-2)/4 of the 256-wide tile.  TMEM loads are software-pipelined one chunk ahead.
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        uint32_t tcount = 0;
        // bias of a tile's 256 columns lives in shared memory (per-chunk global loads were the top
        // stall of the epilogue), double buffered by tile parity.  The values of tile i + 1 are
        // fetched into a register while tile i is processed and stored at its end, so the global
        // load latency is off the per-tile critical path; one 256-thread named barrier per tile.
        const int et = threadIdx.x - 64;                             // 0..255 among the epilogue warps
        const uint32_t sbias_u32 = smem_u32(sbias);
        if (p.bias && (int)blockIdx.x < num_tiles) {
            int mt, nt;
            tmap_((int)blockIdx.x / ksplit, mt, nt);
            const int n0 = nt * BN;
            sbias[et] = (n0 + et < p.N) ? __ldg(p.bias + n0 + et) : 0.f;
        }
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            int mt, nt;
            const int ks = tile % ksplit;                            // split-K part (out_mode 0): bias / residual in part 0
            tmap_(tile / ksplit, mt, nt);
            const int m0 = mt * BM, n0 = nt * BN;
            const uint32_t buf = tcount & 1;
            const int m = m0 + row;
            const uint32_t sb = sbias_u32 + buf * BN * 4;
            float next_bias = 0.f;
            if (p.bias) {
                asm volatile("bar.sync 1, 256;" ::: "memory");      // this tile's bias visible; previous tile's readers done
                const int ntile = tile + gridDim.x;
                if (ntile < num_tiles) {
                    int mt2, nt2;
                    tmap_(ntile / ksplit, mt2, nt2);
                    const int nn = nt2 * BN + et;
                    if (nn < p.N) next_bias = __ldg(p.bias + nn);
                }
            }
            const bool tr = lane == 0 && quarter == 0;
            const int trole = 2 + chalf;
            if (tr) GEMM_TR(trole, tcount, 0);
            mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
            tcgen05_fence_after();
            if (tr) GEMM_TR(trole, tcount, 1);
            int kv_b = 0, kv_key = 0;
            if ((OM == 2 || OM == 4) && m < p.M) { kv_b = m / p.kv_Lk; kv_key = m - kv_b * p.kv_Lk; }
            // staged output (out_mode 2): decided per warp here, where all 32 lanes are still together
            const int key0w = __shfl_sync(0xffffffffu, kv_key, 0);
            const bool slab = (OM == 4 ? p.tma_out : p.staged) && (m0 + quarter * 32 + 32 <= p.M) && (key0w + 32 <= p.kv_Lk);
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + chalf * 128;
            uint32_t rbuf[2][32];
            tmem_ld_x32(t_addr, rbuf[0]);
            if (OM == 2 && p.staged) {           // this warp's four store slabs: the previous tile's TMA stores have read them
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
#pragma unroll 1
            for (int cpair = 0; cpair < 2; ++cpair) {
            int slab_row[2] = {-1, -1};
            uint32_t slab_src[2] = {0u, 0u};
            int slab_col[2] = {0, 0};                      // out_mode 3: column of the hi half (lo = + N)
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) {
                const int cc = cpair * 2 + ci;
                const int c = chalf * 4 + cc;
                uint32_t* r = rbuf[ci];
                // residual of this chunk in the coalesced (transposed) pattern, requested before the TMEM wait so
                // that its L2 latency overlaps the accumulator load and the bias / activation math
                float4 res4[8];
                const bool coal = p.coalesce && (n0 + c * 32 + 32 <= p.N) && OM != 2;
                const bool use_res = p.residual && ks == 0;
                if (coal && use_res) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int mr = m0 + quarter * 32 + (q & 3) * 8 + (lane >> 2);
                        const int col = n0 + c * 32 + (q >> 2) * 16 + (lane & 3) * 4;
                        const int rr = (OM == 5) ? mr % p.kv_Lk : mr;          // mode 5: residual rows repeat with period kv_Lk
                        res4[q] = (mr < p.M) ? __ldg(reinterpret_cast<const float4*>(p.residual + (int64_t)rr * p.ldc + col))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                tmem_ld_wait();
                if (tr) GEMM_TR(trole, tcount, 2 + cc);
                if (cc + 1 < 4) tmem_ld_x32(t_addr + (cc + 1) * 32, rbuf[ci ^ 1]);
                const int nb = n0 + c * 32;
                if (nb >= p.N) continue;
                if (p.dbg == 2) continue;                            // diagnostic: TMEM loads only (the MMA-bound floor)
                const bool full = (nb + 32 <= p.N);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (p.bias && ks == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 t = ld_shared_f32x4(sb + (c * 32 + j) * 4);
                        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                    }
                }
                if (p.act == PETR_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if constexpr (OM == 7) {                               // PETRv2 SELayer: gate[m,n] * sigmoid(acc + bias)
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
                }
                if (p.alpha != 1.f) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
                }
                if constexpr (OM == 3 || OM == 4) {
                    const bool fast = (OM == 3) ? (p.tma_out != 0) : slab;              // warp-uniform
                    if (fast) {
                        // hi / lo halves of this 32 x 32 chunk -> slabs 2 ci and 2 ci + 1 of the warp (64-byte TMA swizzle)
                        if (ci == 0) {                          // the pair's four slabs: previous pair's stores have read them
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            __syncwarp();
                        }
                        const int s0 = ci * 2;
                        uint8_t* swh = smem + 3 * STAGE_BYTES + ((warp - 2) * 3 + s0) * 2048;
                        uint8_t* swl = (s0 + 1 < 3) ? smem + 3 * STAGE_BYTES + ((warp - 2) * 3 + s0 + 1) * 2048
                                                    : stage_out + (warp - 2) * 2048;
                        const uint32_t ah = smem_u32(swh) + lane * 64, al = smem_u32(swl) + lane * 64;
                        const int xh = (ah >> 7) & 3, xl = (al >> 7) & 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 hq, lq;
                            split_bf16x2(v[8 * j], v[8 * j + 1], hq.x, lq.x); split_bf16x2(v[8 * j + 2], v[8 * j + 3], hq.y, lq.y);
                            split_bf16x2(v[8 * j + 4], v[8 * j + 5], hq.z, lq.z); split_bf16x2(v[8 * j + 6], v[8 * j + 7], hq.w, lq.w);
                            st_shared_v4(ah + ((j ^ xh) << 4), hq.x, hq.y, hq.z, hq.w);
                            st_shared_v4(al + ((j ^ xl) << 4), lq.x, lq.y, lq.z, lq.w);
                        }
                        slab_src[ci] = smem_u32(swh);
                        slab_col[ci] = (int)smem_u32(swl);       // (re-used as the lo slab's address)
                        if constexpr (OM == 3) {
                            slab_row[ci] = nb;                   // column of the hi half; row = m0 + 32 quarter
                        } else {
                            const int l = nb / p.kv_E, h = (nb % p.kv_E) >> 5;
                            slab_row[ci] = (int)((((int64_t)l * p.kv_B + kv_b) * (p.kv_E >> 5) + h) * p.kv_Lk + key0w);
                        }
                        continue;
                    }
                }
                if (coal) {
                    // Coalesced stores: a thread owns one ROW of the accumulator (TMEM lane), so direct stores put
                    // 16 bytes into each of 32 different lines per instruction (measured: the epilogue, not the
                    // MMA, paced every GEMM -- 2 700 clk per 32-column chunk for bf16 rows, 4 300 for fp32 rows).
                    // The warp's 32 x 32 block is transposed through a 2 KB shared-memory slab (16 columns at a
                    // time, XOR-swizzled: conflict-free both ways) so that each store / residual-load instruction
                    // covers 8 rows x 64 contiguous bytes.
                    const uint32_t sw = smem_u32(stage_out) + (warp - 2) * 2048;
                    const int mw = m0 + quarter * 32;
                    const int l = nb / p.kv_E, h = (nb % p.kv_E) >> 5;           // out_mode 4: chunk = one head
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            st_shared_v4(sw + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
                                         __float_as_uint(v[16 * hh + 4 * j]), __float_as_uint(v[16 * hh + 4 * j + 1]),
                                         __float_as_uint(v[16 * hh + 4 * j + 2]), __float_as_uint(v[16 * hh + 4 * j + 3]));
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rl = i * 8 + (lane >> 2), piece = lane & 3;
                            float4 x = ld_shared_f32x4(sw + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
                            const int mr = mw + rl;
                            if (mr >= p.M) continue;
                            const int col = nb + hh * 16 + piece * 4;
                            if (OM == 5) {                                    // bf16 without, fp32 with the residual
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)mr * p.N + col) =
                                    make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
                                const float4 q4 = res4[hh * 4 + i];
                                x.x += q4.x; x.y += q4.y; x.z += q4.z; x.w += q4.w;
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C2) + (int64_t)mr * p.ldc + col) = x;
                                continue;
                            }
                            if constexpr (OM == 7) {
                                const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gate + (int64_t)mr * p.ldc + col));
                                x.x *= g4.x; x.y *= g4.y; x.z *= g4.z; x.w *= g4.w;
                            }
                            if (use_res) {
                                const float4 q4 = res4[hh * 4 + i];
                                x.x += q4.x; x.y += q4.y; x.z += q4.z; x.w += q4.w;
                            }
                            if (OM == 0 || OM == 7) {
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (int64_t)ks * p.M * p.ldc +
                                                           (int64_t)mr * p.ldc + col) = x;
                            } else if (OM == 6) {                             // fp32 [hi | lo] tf32 split rows
                                float4 hi4, lo4;
                                hi4.x = tf32_rna(x.x); hi4.y = tf32_rna(x.y); hi4.z = tf32_rna(x.z); hi4.w = tf32_rna(x.w);
                                lo4.x = x.x - hi4.x; lo4.y = x.y - hi4.y; lo4.z = x.z - hi4.z; lo4.w = x.w - hi4.w;
                                float* cp = reinterpret_cast<float*>(p.C) + (int64_t)mr * (2 * p.N) + col;
                                *reinterpret_cast<float4*>(cp) = hi4;
                                *reinterpret_cast<float4*>(cp + p.N) = lo4;
                            } else if (OM == 1) {
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)mr * p.ldc + col) =
                                    make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
                            } else {
                                uint2 hi, lo;
                                split_bf16x2(x.x, x.y, hi.x, lo.x);
                                split_bf16x2(x.z, x.w, hi.y, lo.y);
                                if (OM == 3) {
                                    if (p.C2) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C2) + (int64_t)mr * p.ldc + col) = x;
                                    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)mr * (2 * p.N) + col;
                                    *reinterpret_cast<uint2*>(cp) = hi;
                                    *reinterpret_cast<uint2*>(cp + p.N) = lo;
                                } else {                                      // out_mode 4
                                    const int bb = mr / p.kv_Lk, key = mr - bb * p.kv_Lk;
                                    const int64_t off = ((((int64_t)l * p.kv_B + bb) * (p.kv_E >> 5) + h) * p.kv_Lk + key) * 32 + (col & 31);
                                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off) = hi;
                                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C2) + off) = lo;
                                }
                            }
                        }
                    }
                    continue;
                }
                if (m >= p.M) continue;
                if constexpr (OM == 7) {
                    const float* gp = p.gate + (int64_t)m * p.ldc + nb;
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (nb + j < p.N) v[j] *= gp[j];
                }
                if (use_res) {
                    const float* rp = p.residual + (int64_t)m * p.ldc + nb;
                    if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 t = *reinterpret_cast<const float4*>(rp + j);
                            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (nb + j < p.N) v[j] += rp[j];
                    }
                }
                if (OM == 6) {
                    float* cp = reinterpret_cast<float*>(p.C) + (int64_t)m * (2 * p.N) + nb;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < p.N) { const float h = tf32_rna(v[j]); cp[j] = h; cp[p.N + j] = v[j] - h; }
                }
                if (OM == 0 || OM == 7 || (OM == 3 && p.C2)) {
                    float* cp = reinterpret_cast<float*>(OM == 3 ? p.C2 : p.C) + (OM == 0 ? (int64_t)ks * p.M * p.ldc : 0) +
                                (int64_t)m * p.ldc + nb;
                    if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (nb + j < p.N) cp[j] = v[j];
                    }
                }
                if (OM == 1) {
                    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)m * p.ldc + nb;
                    if (full && ((p.ldc & 7) == 0)) store_bf16x32(cp, v);
                    else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (nb + j < p.N) cp[j] = __float2bfloat16_rn(v[j]);
                    }
                } else if (OM == 2) {
                    // head-split layout [n / E][b][head][key][32]; a 32-column chunk is one head
                    const int l = nb / p.kv_E, h = (nb % p.kv_E) >> 5;
                    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) +
                        ((((int64_t)l * p.kv_B + kv_b) * (p.kv_E >> 5) + h) * p.kv_Lk + kv_key) * 32;
                    // the warp's 32 rows are 32 consecutive keys of one
                    // (layer, batch, head) = 2 KB contiguous in global memory whenever they do not
                    // straddle a batch boundary or the end of M.  Per-thread stores put 16 B into each
                    // of 32 sectors per instruction (ncu: 73.7 M sector writes for 1.18 GB); staged,
                    // one elected lane issues a single 2 KB cp.async.bulk.
                    if (slab) {
                        // slab cc of this warp (three in the freed operand stage, the fourth in stage_out)
                        uint8_t* sw = (cc < 3) ? smem + 3 * STAGE_BYTES + ((warp - 2) * 3 + cc) * 2048
                                               : stage_out + (warp - 2) * 2048;
                        const uint32_t sa = smem_u32(sw) + lane * 64;
                        const int x = (sa >> 7) & 3;                              // SWIZZLE_64B: 16-byte chunk index ^= address bits [7:8]
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            st_shared_v4(sa + ((j ^ x) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                         pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                        slab_row[ci] = (int)((((int64_t)l * p.kv_B + kv_b) * (p.kv_E >> 5) + h) * p.kv_Lk + key0w);
                        slab_src[ci] = smem_u32(sw);
                    } else if (p.dbg != 1) {
                        slab_row[ci] = -1;
                        store_bf16x32(cp, v);
                    }
                } else if (OM == 4) {
                    // head-split [hi | lo] pair: C = hi array, C2 = lo array, both [n / E][b][head][key][32]
                    const int l = nb / p.kv_E, h = (nb % p.kv_E) >> 5;
                    const int64_t off = ((((int64_t)l * p.kv_B + kv_b) * (p.kv_E >> 5) + h) * p.kv_Lk + kv_key) * 32;
                    store_split_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.C) + off,
                                        reinterpret_cast<__nv_bfloat16*>(p.C2) + off, v);
                } else if (OM == 3) {
                    // bf16 split [hi | lo] for the bf16x3 GEMMs: row m, hi at [nb], lo at [N + nb]
                    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)m * (2 * p.N) + nb;
                    store_split_bf16x32(cp, cp + p.N, v);
                }
            }
            if ((OM == 2 || OM == 3 || OM == 4) && (slab_row[0] >= 0 || slab_row[1] >= 0)) {   // warp-uniform: one proxy fence per two chunks
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && p.dbg != 1) {
#pragma unroll
                    for (int ci = 0; ci < 2; ++ci) {
                        if (slab_row[ci] < 0) continue;
                        if constexpr (OM == 2) {
                            tma_store_2d(&tmC, slab_src[ci], 0, slab_row[ci]);
                        } else if constexpr (OM == 3) {
                            tma_store_2d(&tmC, slab_src[ci], slab_row[ci], m0 + quarter * 32);
                            tma_store_2d(&tmC, (uint32_t)slab_col[ci], p.N + slab_row[ci], m0 + quarter * 32);
                        } else {
                            tma_store_2d(&tmC, slab_src[ci], 0, slab_row[ci]);
                            tma_store_2d(&tmD, (uint32_t)slab_col[ci], 0, slab_row[ci]);
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            }
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[buf]);
            if (tr) GEMM_TR(trole, tcount, 6);
            if (p.bias) sbias[(buf ^ 1) * BN + et] = next_bias;     // read after the next tile's barrier
        }
        if ((p.staged || p.tma_out) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// fp32 NCHW feature maps (G, C, hw) -> bf16 token-major (G*hw, C): A operand of input_proj.
// One CTA moves a 32-pixel x 64-channel block: every global read is a 128-byte row segment (32 pixels of one
// channel), every global write a 128-byte row segment (64 bf16 channels of one pixel); 8 independent loads per
// thread are in flight before the first use (HBM-bound: 4 B in + 2 B out per element).
__device__ __forceinline__ float ld_as_f32(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_f32(const __nv_bfloat16* p) {
    return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

// InT = float, or __nv_bfloat16 for feature maps handed over in half precision (an AMP backbone / the bf16 wire
// format of bench.py's e2e_bf16_features: half the host -> device bytes, bit-identical tokens).
template <typename InT>
__global__ void __launch_bounds__(256)
nchw_to_tokens_bf16_kernel(const InT* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int hw) {
    __shared__ float tile[64][33];
    const int g = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;          // 8 warps
    const InT* xg = x + (int64_t)g * C * hw;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + w * 8 + i, pp = p0 + lane;
        v[i] = (c < C && pp < hw) ? ld_as_f32(xg + (int64_t)c * hw + pp) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[w * 8 + i][lane] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pl = w * 4 + i, pp = p0 + pl, c = c0 + 2 * lane;
        if (pp < hw && c < C) {
            __nv_bfloat16* o = out + ((int64_t)g * hw + pp) * C + c;
            if (c + 1 < C) *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(tile[2 * lane][pl], tile[2 * lane + 1][pl]);
            else *o = __float2bfloat16_rn(tile[2 * lane][pl]);
        }
    }
}

// fp32 rows (R, E) (+ add[r % add_rows]) -> fp32 [hi | lo] rows (R, 2E): hi = tf32(x), lo = x - hi (3xTF32 operand)
__global__ void split_rows_tf32_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ add, int add_rows,
                                       float* __restrict__ out, int64_t total, int E) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t r = i / E;
    const int c = (int)(i - r * E);
    float v = X[r * ldx + c];
    if (add) v += add[(r % add_rows) * (int64_t)E + c];
    const float h = tf32_rna(v);
    out[r * 2 * E + c] = h;
    out[r * 2 * E + E + c] = v - h;
}

}  // namespace
}  // namespace petr

using namespace petr;

extern "C" int petr_split_rows_tf32(const float* X, int ldx, const float* add, int add_rows, float* out, int rows, int E,
                                    void* stream) {
    PETR_REQUIRE(X && out && rows > 0 && E > 0 && ldx >= E, PETR_E_BADARG, "petr_split_rows_tf32: bad argument");
    PETR_REQUIRE(!add || add_rows > 0, PETR_E_BADARG, "petr_split_rows_tf32: add needs add_rows > 0");
    const int64_t total = (int64_t)rows * E;
    split_rows_tf32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, add, add_rows, out, total, E);
    return after_launch("petr_split_rows_tf32");
}

extern "C" int petr_gemm_tf32x3_tc(const float* A, int lda, int a_k_wrap, const float* W, int ldw, const float* bias,
                                   const float* residual, int act, float* C, int ldc, int out_mode, int M, int N, int K,
                                   void* stream) {
    PETR_REQUIRE(A && W && C, PETR_E_BADARG, "petr_gemm_tf32x3_tc: null pointer");
    PETR_REQUIRE(M > 0 && N > 0 && K > 0, PETR_E_BADARG, "petr_gemm_tf32x3_tc: bad dimension");
    PETR_REQUIRE(K % 32 == 0, PETR_E_SHAPE, "petr_gemm_tf32x3_tc: K must be a multiple of 32, got %d", K);
    if (a_k_wrap <= 0) a_k_wrap = K;
    PETR_REQUIRE(a_k_wrap % 32 == 0 && a_k_wrap <= K && K - a_k_wrap <= a_k_wrap, PETR_E_SHAPE,
                 "petr_gemm_tf32x3_tc: a_k_wrap must be a multiple of 32 with K/2 <= a_k_wrap <= K");
    PETR_REQUIRE(lda >= a_k_wrap && ldw >= K && (lda % 4) == 0 && (ldw % 4) == 0 && aligned16(A) && aligned16(W),
                 PETR_E_ALIGN, "petr_gemm_tf32x3_tc: A/W rows must be 16-byte aligned");
    PETR_REQUIRE(act == PETR_ACT_NONE || act == PETR_ACT_RELU, PETR_E_BADARG, "petr_gemm_tf32x3_tc: bad act");
    PETR_REQUIRE(out_mode == 0 || out_mode == 6, PETR_E_BADARG, "petr_gemm_tf32x3_tc: out_mode must be 0 (fp32) or 6 (split)");
    PETR_REQUIRE((N % 4) == 0 && aligned16(C) && (!residual || aligned16(residual)), PETR_E_ALIGN,
                 "petr_gemm_tf32x3_tc: N %% 4 == 0 and 16-byte aligned C / residual required");
    if (out_mode == 0 || residual) PETR_REQUIRE(ldc >= N && (ldc % 4) == 0, PETR_E_BADARG, "petr_gemm_tf32x3_tc: bad ldc");
    EpiParams p{};
    p.bias = bias; p.residual = residual; p.gate = nullptr; p.act = act; p.alpha = 1.f;
    p.C = C; p.C2 = nullptr; p.ldc = ldc; p.out_mode = out_mode;
    p.M = M; p.N = N; p.K = K; p.a_k_wrap = a_k_wrap; p.kv_E = 256;
    p.trace = nullptr;
    p.coalesce = 1;
    p.ksplit = 1;
    CUtensorMap tmA, tmB;
    uint64_t dA[2] = {(uint64_t)a_k_wrap, (uint64_t)M}, sA[1] = {(uint64_t)lda * 4};
    uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)ldw * 4};
    uint32_t boxA[2] = {32, BM}, boxB[2] = {32, BN};
    int rc = tmap::encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dA, sA, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = tmap::encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dB, sB, boxB, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    typedef void (*gemm_fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, EpiParams);
    const gemm_fn fn = out_mode == 0 ? (gemm_fn)gemm_bf16_tc_kernel<0, true> : (gemm_fn)gemm_bf16_tc_kernel<6, true>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) return fail((int)e, "petr_gemm_tf32x3_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = (N + BN - 1) / BN;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = p.tiles_m * p.tiles_n;
    fn<<<tiles < sms ? tiles : sms, GEMM_THREADS, GEMM_SMEM, (cudaStream_t)stream>>>(tmA, tmB, tmA, tmA, p);
    return after_launch("petr_gemm_tf32x3_tc");
}

extern "C" int petr_nchw_to_tokens_bf16(const void* x, int in_dtype, void* out, int G, int C, int hw, void* stream) {
    PETR_REQUIRE(x && out, PETR_E_BADARG, "petr_nchw_to_tokens_bf16: null pointer");
    PETR_REQUIRE(G > 0 && C > 0 && hw > 0 && G <= 65535, PETR_E_BADARG, "petr_nchw_to_tokens_bf16: bad dimension");
    PETR_REQUIRE((C % 2) == 0, PETR_E_SHAPE, "petr_nchw_to_tokens_bf16: C must be even, got %d", C);
    dim3 grid((hw + 31) / 32, (C + 63) / 64, G);
    if (in_dtype == PETR_DTYPE_F32)
        nchw_to_tokens_bf16_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (__nv_bfloat16*)out, C, hw);
    else if (in_dtype == PETR_DTYPE_BF16)
        nchw_to_tokens_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x,
                                                                                          (__nv_bfloat16*)out, C, hw);
    else
        return fail(PETR_E_BADARG, "petr_nchw_to_tokens_bf16: bad in_dtype %d", in_dtype);
    return after_launch("petr_nchw_to_tokens_bf16");
}

extern "C" int petr_gemm_bf16_tc(const void* A, int lda, int a_k_wrap, const void* W, int ldw,
                                 const float* bias, const float* residual, const float* gate, int act, float alpha,
                                 void* C, void* C2, int ldc, int out_mode, int M, int N, int K,
                                 int kv_Lk, int ksplit, void* stream) {
    PETR_REQUIRE(A && W && C, PETR_E_BADARG, "petr_gemm_bf16_tc: null pointer");
    PETR_REQUIRE(M > 0 && N > 0 && K > 0, PETR_E_BADARG, "petr_gemm_bf16_tc: bad dimension");
    PETR_REQUIRE(K % BK == 0, PETR_E_SHAPE, "petr_gemm_bf16_tc: K must be a multiple of 64, got %d", K);
    if (a_k_wrap <= 0) a_k_wrap = K;
    PETR_REQUIRE(a_k_wrap % BK == 0 && a_k_wrap <= K && K - a_k_wrap <= a_k_wrap, PETR_E_SHAPE,
                 "petr_gemm_bf16_tc: a_k_wrap must be a multiple of 64 with K/2 <= a_k_wrap <= K");
    PETR_REQUIRE(lda >= a_k_wrap && ldw >= K && (lda % 8) == 0 && (ldw % 8) == 0 && aligned16(A) && aligned16(W),
                 PETR_E_ALIGN, "petr_gemm_bf16_tc: A/W rows must be 16-byte aligned");
    PETR_REQUIRE(act == PETR_ACT_NONE || act == PETR_ACT_RELU || act == PETR_ACT_SIGMOID_GATE, PETR_E_BADARG,
                 "petr_gemm_bf16_tc: bad act");
    PETR_REQUIRE((act == PETR_ACT_SIGMOID_GATE) == (gate != nullptr), PETR_E_BADARG,
                 "petr_gemm_bf16_tc: PETR_ACT_SIGMOID_GATE needs a gate array (and only it)");
    if (gate) PETR_REQUIRE(out_mode == 0 && aligned16(gate) && ldc >= N, PETR_E_SHAPE,
                           "petr_gemm_bf16_tc: a gate needs out_mode 0 and a 16-byte aligned [M, ldc] array");
    PETR_REQUIRE(out_mode >= 0 && out_mode <= 5, PETR_E_BADARG, "petr_gemm_bf16_tc: bad out_mode");
    EpiParams p{};
    p.bias = bias; p.residual = residual; p.gate = gate; p.act = act; p.alpha = alpha;
    p.C = C; p.C2 = C2; p.ldc = ldc; p.out_mode = out_mode;
    p.M = M; p.N = N; p.K = K; p.a_k_wrap = a_k_wrap;
    p.staged = (out_mode == 2 && tuning::gemm_staged()) ? 1 : 0;
    p.dbg = tuning::gemm_debug();
    if (ksplit < 1) ksplit = 1;
    PETR_REQUIRE(ksplit == 1 || (out_mode == 0 && !gate && ksplit <= K / BK), PETR_E_BADARG,
                 "petr_gemm_bf16_tc: ksplit > 1 needs out_mode 0 without a gate and at most K / 64 parts");
    p.ksplit = ksplit;
    p.trace = (long long*)tuning::debug_trace();
    p.coalesce = (tuning::gemm_coalesce() && aligned16(C) && (!C2 || aligned16(C2)) && (!residual || aligned16(residual)) &&
                  (N % 4) == 0 && ((out_mode == 4) || (ldc % 4) == 0)) ? 1 : 0;
    if (out_mode != 2 && out_mode != 4) p.kv_E = 256;
    if (out_mode == 2 || out_mode == 4) {
        if (out_mode == 4) PETR_REQUIRE(C2 && aligned16(C2), PETR_E_BADARG, "petr_gemm_bf16_tc: out_mode 4 needs C2 (lo array)");
        PETR_REQUIRE(kv_Lk > 0 && M % kv_Lk == 0 && N % 256 == 0 && !residual, PETR_E_SHAPE,
                     "petr_gemm_bf16_tc: head-split output needs M %% Lk == 0, N %% 256 == 0, no residual");
        p.kv_Lk = kv_Lk; p.kv_B = M / kv_Lk; p.kv_E = 256;
        PETR_REQUIRE(aligned16(C), PETR_E_ALIGN, "petr_gemm_bf16_tc: C must be 16-byte aligned");
    } else if (out_mode == 3) {
        PETR_REQUIRE(N % 32 == 0 && aligned16(C), PETR_E_SHAPE, "petr_gemm_bf16_tc: split output needs N %% 32 == 0");
        if (C2 || residual) PETR_REQUIRE(ldc >= N, PETR_E_BADARG, "petr_gemm_bf16_tc: ldc < N");
    } else {
        PETR_REQUIRE(ldc >= N, PETR_E_BADARG, "petr_gemm_bf16_tc: ldc < N");
    }
    CUtensorMap tmA, tmB;
    uint64_t dA[2] = {(uint64_t)a_k_wrap, (uint64_t)M}, sA[1] = {(uint64_t)lda * 2};
    uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)ldw * 2};
    uint32_t boxA[2] = {BK, BM}, boxB[2] = {BK, BN};
    int rc = tmap::encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dA, sA, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = tmap::encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dB, sB, boxB, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    CUtensorMap tmC = tmA, tmD = tmA;              // output maps of the TMA-store epilogues (out_modes 2, 3, 4)
    if (p.staged || (out_mode == 4 && tuning::gemm_staged())) {
        uint64_t dC[2] = {32, (uint64_t)(N / 256) * p.kv_B * 8 * (uint64_t)p.kv_Lk}, sC[1] = {64};
        uint32_t boxC[2] = {32, 32};
        rc = tmap::encode(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, dC, sC, boxC, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        if (out_mode == 4) {
            rc = tmap::encode(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C2, dC, sC, boxC, CU_TENSOR_MAP_SWIZZLE_64B);
            if (rc) return rc;
            p.tma_out = 1;
        }
    } else if (out_mode == 3 && !C2 && !residual && tuning::gemm_staged() && p.coalesce) {
        uint64_t dC[2] = {(uint64_t)2 * N, (uint64_t)M}, sC[1] = {(uint64_t)2 * N * 2};
        uint32_t boxC[2] = {32, 32};
        rc = tmap::encode(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, dC, sC, boxC, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        p.tma_out = 1;
    }
    typedef void (*gemm_fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, EpiParams);
    // index 6: out_mode 0 with the sigmoid-gate epilogue -- its own instantiation (kernel-internal mode 7), so that the 32
    // exponentials + divisions per chunk do not bloat the other kernels (they doubled every instantiation's SASS, 34 ->
    // 63 KB against a 32 KB instruction cache, when the gate was a run-time branch)
    static const gemm_fn fns[7] = {gemm_bf16_tc_kernel<0>, gemm_bf16_tc_kernel<1>, gemm_bf16_tc_kernel<2>,
                                   gemm_bf16_tc_kernel<3>, gemm_bf16_tc_kernel<4>, gemm_bf16_tc_kernel<5>,
                                   gemm_bf16_tc_kernel<7>};
    if (out_mode == 5)
        PETR_REQUIRE(p.coalesce && C2 && residual && kv_Lk > 0 && N % 32 == 0 && ldc >= N && aligned16(C), PETR_E_SHAPE,
                     "petr_gemm_bf16_tc: out_mode 5 needs C2, a residual with period kv_Lk, N %% 32 == 0 and 16-byte aligned bases");
    if (out_mode == 5) p.kv_Lk = kv_Lk;
    const gemm_fn fn = fns[gate ? 6 : out_mode];
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) return fail((int)e, "petr_gemm_bf16_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = (N + BN - 1) / BN;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int tiles = p.tiles_m * p.tiles_n * ksplit;
    int grid = tiles < sms ? tiles : sms;
    p.band = (out_mode == 2 && p.tiles_n > 1 && tuning::gemm_band()) ? grid : 0;      // measured slower (0.509 vs 0.466 ms): off by default
    e = launch_pdl(fn, dim3(grid), dim3(GEMM_THREADS), GEMM_SMEM, (cudaStream_t)stream, tmA, tmB, tmC, tmD, p);
    if (e != cudaSuccess) return fail((int)e, "petr_gemm_bf16_tc: launch: %s", cudaGetErrorString(e));
    return after_launch("petr_gemm_bf16_tc");
}
