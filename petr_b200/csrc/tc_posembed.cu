// Fused 3D position embedding (PETRHead.position_embeding, petr_head.py:286-334) on tcgen05:
//
//   frustum (u*d, v*d, d, 1)  ->  img2lidar 4x4  ->  position_range normalise  ->  inverse_sigmoid
//   ->  Conv1x1(192 -> 1024) + ReLU  ->  Conv1x1(1024 -> 256)  (+ bias, + optional residuals)
//
// in ONE persistent kernel (one CTA per SM looping over 128-token tiles): the (B,N,W,H,D,4) coordinate
// tensor, the 192-channel MLP input and the 1024-wide hidden activation never touch HBM, and the hidden
// activation does not even touch shared memory.  Warp roles (14 warps):
//   0       TMA: streams W1 / W2 blocks (16 KB) through a 6-slot mbarrier ring, free-running across tiles
//   1       tcgen05.mma issuer (all 128 x 128 x 16; the whole warp walks the loop, the elected lane issues)
//   2-5     feature producers, one thread per token: the img2lidar matrices of the tile's views are staged in
//           shared memory with one vectorised load per row, then the 64 depth bins of the pixel are evaluated
//           and the 192 bf16 features written straight into the UMMA K-major 128B-swizzled A operand.  The A
//           tile is double buffered: tile i + 1 is generated while the MMAs of tile i run
//   6-9     hidden-chunk epilogue: GEMM1 runs in 8 chunks of 128 hidden units into a double-buffered TMEM
//           accumulator; these warps apply bias + ReLU, convert to bf16 and write the chunk back into the SAME
//           tensor-memory columns (first 64 of the 128), from where GEMM2 takes it as its A operand (TS MMA)
//   10-13   output epilogue: O (128 x 256, TMEM) + b2 (+ residual, + broadcast residual2) -> fp32 / bf16 tokens,
//           transposed through small shared-memory slabs so that every load / store instruction covers
//           8 rows x 64 contiguous bytes; overlaps the first two hidden chunks of the next tile
// The bf16 path rounds the features to bf16 anyway, so the coordinate arithmetic uses fast fp32
// (FMA contraction, __fdividef, __logf); the bit-exact coordinates live in petr_frustum_features.
#include "common.cuh"
#include "tc_common.cuh"
#include <math.h>


namespace petr {
namespace {

using namespace tc;

constexpr int TM = 128;                 // tokens per tile
constexpr int K1 = 192;                 // 3 * depth_num
constexpr int HID = 1024, EOUT = 256;
constexpr int HC = 128;                 // hidden units per chunk
constexpr int NCH = HID / HC;           // 8
constexpr int A1_BYTES = TM * K1 * 2;                   // 48 KB: 3 K-atoms of 16 KB
constexpr int WST = 6;                                  // weight ring: 6 slots of [128 rows x 64 k] = 16 KB
constexpr int WST_BYTES = 128 * 64 * 2;
constexpr int MAXV = 16;                                // views whose matrices are staged per tile
constexpr int OFF_A1 = 0;                               // 2 buffers
constexpr int OFF_WR = OFF_A1 + 2 * A1_BYTES;           // weight ring (W1 k-blocks and W2 half k-blocks, in MMA order)
constexpr int OFF_SLAB = OFF_WR + WST * WST_BYTES;      // 4 output-epilogue warps x 2 KB transposing slab
constexpr int OFF_MAT = OFF_SLAB + 4 * 2048;            // 2 x MAXV x 12 floats (img2lidar rows 0-2)
constexpr int OFF_B1 = OFF_MAT + 2 * MAXV * 48;         // 1024 floats
constexpr int OFF_B2 = OFF_B1 + HID * 4;                // 256 floats
constexpr int OFF_BAR = OFF_B2 + EOUT * 4;
constexpr int PE_SMEM = OFF_BAR + 256 + 1024;
constexpr int PE_THREADS = 14 * 32;
constexpr int TM_H = 0;                 // H chunk accumulators at columns [0,128) and [128,256); bf16 copy in the first 64
constexpr int TM_O = 256;               // output accumulator, 256 columns

struct PeParams {
    const float* img2lidar;             // [B*N, 16]
    int B, N, H, W, pad_h, pad_w, lid;
    float depth_start, bin_size;
    float scale[3], shift[3];           // x_norm = X * scale + shift  (= (X - lo) / span)
    const float* b1; const float* b2; const float* residual; float* out;
    const float* residual2; int res2_rows; __nv_bfloat16* out_bf16;   // + residual2[t % res2_rows]; optional bf16 store
    int64_t T; int num_tiles;
    long long* trace;                   // diagnostic clock64 stamps of CTA 0 (petr_debug_set_trace), normally null
};

// role 0 MMA [chunk g][0: h_free ok, 1: GEMM1 issued, 2: p_ready ok (chunk g-1), 3: GEMM2 issued]; role 1 producer
// [tile][0: a1_empty ok, 1: done]; role 2 hidden epilogue [g][0: h_full ok, 1: done]; role 3 output epilogue [tile]
// [0: o_full ok, 1: o_free arrived, 2: done]
#define PE_TR(role, t, ev) do { if (p.trace && blockIdx.x == 0 && (t) < 64) p.trace[((role) * 64 + (t)) * 8 + (ev)] = clock64(); } while (0)


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
    uint64_t* a1_full = bars;             // [2], 128 producers
    uint64_t* a1_empty = a1_full + 2;     // [2], MMA commit
    uint64_t* w_full = a1_empty + 2;      // [WST]
    uint64_t* w_empty = w_full + WST;     // [WST]
    uint64_t* h_full = w_empty + WST;     // [2], MMA commit: GEMM1 chunk in TMEM
    uint64_t* p_ready = h_full + 2;       // [2], 128: bf16 hidden chunk written back to TMEM
    uint64_t* h_free = p_ready + 2;       // [2], MMA commit: GEMM2 has consumed the chunk
    uint64_t* o_full = h_free + 2;        // MMA commit: tile's output accumulator complete
    uint64_t* o_free = o_full + 1;        // 128: output accumulator read back
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
    float* sb1 = reinterpret_cast<float*>(smem + OFF_B1);
    float* sb2 = reinterpret_cast<float*>(smem + OFF_B2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas
    const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    for (int i = threadIdx.x; i < HID; i += PE_THREADS) sb1[i] = p.b1[i];
    for (int i = threadIdx.x; i < EOUT; i += PE_THREADS) sb2[i] = p.b2[i];
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&a1_full[i], 128); mbar_init(&a1_empty[i], 1);
                mbar_init(&h_full[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&h_free[i], 1);
            }
            for (int i = 0; i < WST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
            mbar_init(o_full, 1); mbar_init(o_free, 128);
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
        // ---------------- TMA: weights through the ring, in the MMA's order, free-running across tiles ----------------
        // per hidden chunk c: 3 k-blocks of W1[c] (GEMM1), then the 4 half k-blocks of W2[:, c-1] (GEMM2)
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
        for (int il = 0; il < my_tiles; ++il) {
            for (int c = 0; c < NCH; ++c) {
                for (int kb = 0; kb < K1 / 64; ++kb) load(&tmW1, kb * 64, c * HC);
                if (c > 0) load_w2(c - 1);
            }
            load_w2(NCH - 1);
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(TM, 128, 1, 0, 0);           // every MMA here is 128 x 128 x 16
        const uint64_t a1d0 = make_smem_desc(smem_u32(smem + OFF_A1), 16, 1024, LAYOUT_SW128);
        const uint64_t wrd = make_smem_desc(smem_u32(smem + OFF_WR), 16, 1024, LAYOUT_SW128);
        uint32_t it = 0;                         // ring slot counter
        uint32_t g = 0;                          // global hidden-chunk counter: TMEM buffer = g & 1
        auto next_stage = [&]() -> uint64_t {
            const int st = it % WST;
            mbar_wait(&w_full[st], (it / WST) & 1);
            tcgen05_fence_after();
            ++it;
            return wrd + (uint64_t)((st * WST_BYTES) >> 4);
        };
        auto release = [&]() { if (leader) umma_commit(&w_empty[(it - 1) % WST]); __syncwarp(); };
        // O[128 x 256] += Hbf16[128 x 128] W2[:, chunk]^T, A operand from tensor memory (8 columns = 16 hidden units)
        auto gemm2 = [&](uint32_t gg, int c, int il) {
            const uint32_t hb = gg & 1;
            mbar_wait(&p_ready[hb], (gg >> 1) & 1);
            if (c == 0) mbar_wait(o_free, (il & 1) ^ 1);                  // previous tile's output has been read back
            tcgen05_fence_after();
            if (leader) PE_TR(0, gg + 1, 2);
            for (int kb = 0; kb < HC / 64; ++kb)
                for (int nh = 0; nh < 2; ++nh) {
                    const uint64_t bd = next_stage();
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ts(tmem_base + TM_O + nh * 128, tmem_base + TM_H + hb * HC + kb * 32 + k * 8,
                                        bd + (uint64_t)((k * 32) >> 4), idesc, (c | kb | k) ? 1u : 0u);
                    }
                    release();
                }
            if (leader) { umma_commit(&h_free[hb]); PE_TR(0, gg + 1, 3); }
            __syncwarp();
        };
        for (int il = 0; il < my_tiles; ++il) {
            const uint32_t ab = il & 1;
            const uint64_t a1d = a1d0 + (uint64_t)((ab * A1_BYTES) >> 4);
            mbar_wait(&a1_full[ab], (il >> 1) & 1);
            tcgen05_fence_after();
            for (int c = 0; c < NCH; ++c, ++g) {
                const uint32_t hb = g & 1;
                mbar_wait(&h_free[hb], ((g >> 1) & 1) ^ 1);               // GEMM2 of chunk g - 2 has consumed this buffer
                tcgen05_fence_after();
                if (leader) PE_TR(0, g, 0);
                for (int kb = 0; kb < K1 / 64; ++kb) {
                    const uint64_t bd = next_stage();
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ss(tmem_base + TM_H + hb * HC, a1d + (uint64_t)((kb * 16384 + k * 32) >> 4),
                                        bd + (uint64_t)((k * 32) >> 4), idesc, (kb | k) ? 1u : 0u);
                    }
                    release();
                }
                if (leader) {
                    umma_commit(&h_full[hb]);
                    if (c == NCH - 1) umma_commit(&a1_empty[ab]);         // the A tile may be regenerated
                    PE_TR(0, g, 1);
                }
                __syncwarp();
                if (c > 0) gemm2(g - 1, c - 1, il);
            }
            gemm2(g - 1, NCH - 1, il);
            if (leader) umma_commit(o_full);
            __syncwarp();
        }
    } else if (warp < 6) {
        // ---------------- feature producers: thread <-> token ----------------
        const int pt = threadIdx.x - 64;                       // 0..127
        const int64_t HW = (int64_t)p.H * p.W;
        for (int il = 0; il < my_tiles; ++il) {
            const uint32_t ab = il & 1;
            const int64_t tile0 = ((int64_t)blockIdx.x + (int64_t)il * gridDim.x) * TM;
            float* smat = reinterpret_cast<float*>(smem + OFF_MAT) + ab * (MAXV * 12);
            mbar_wait(&a1_empty[ab], ((il >> 1) & 1) ^ 1);     // GEMM1 of tile il - 2 has consumed this buffer
            if (pt == 0) PE_TR(1, il, 0);
            // the img2lidar matrices (rows 0-2) of the views this tile touches: one float4 per thread, staged in smem
            const int64_t bn_first = tile0 / HW;
            const int64_t t_last = (tile0 + TM - 1 < p.T ? tile0 + TM - 1 : p.T - 1);
            const int nviews = (int)(t_last / HW - bn_first) + 1;
            const bool staged = nviews <= MAXV;
            if (staged && pt < nviews * 3)
                reinterpret_cast<float4*>(smat)[pt] =
                    __ldg(reinterpret_cast<const float4*>(p.img2lidar + (bn_first + pt / 3) * 16) + (pt % 3));
            asm volatile("bar.sync 2, 128;" ::: "memory");
            const int row = pt;
            const int64_t t = tile0 + row;
            uint8_t* arow = smem + OFF_A1 + ab * A1_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
            const bool live = t < p.T;
            float M[12], u = 0.f, v = 0.f;
#pragma unroll
            for (int i = 0; i < 12; ++i) M[i] = 0.f;
            if (live) {
                const int w = (int)(t % p.W);
                const int h = (int)((t / p.W) % p.H);
                const int64_t bn = t / HW;
                float4 m0, m1, m2;
                if (staged) {
                    const float4* mp = reinterpret_cast<const float4*>(smat) + (int)(bn - bn_first) * 3;
                    m0 = mp[0]; m1 = mp[1]; m2 = mp[2];
                } else {
                    const float4* mp = reinterpret_cast<const float4*>(p.img2lidar + bn * 16);
                    m0 = __ldg(mp); m1 = __ldg(mp + 1); m2 = __ldg(mp + 2);
                }
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
            mbar_arrive(&a1_full[ab]);
            if (pt == 0) PE_TR(1, il, 1);
        }
    } else if (warp < 10) {
        // ---------------- hidden-chunk epilogue: bias + ReLU -> bf16 -> back into tensor memory ----------------
        const int quarter = warp & 3;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t sb1_u32 = smem_u32(sb1);
        const uint32_t total = (uint32_t)my_tiles * NCH;
        for (uint32_t g = 0; g < total; ++g) {
            const uint32_t hb = g & 1;
            const int c = (int)(g % NCH);
            mbar_wait(&h_full[hb], (g >> 1) & 1);
            tcgen05_fence_after();
            if (threadIdx.x == 6 * 32) PE_TR(2, g, 0);
            const uint32_t ha = tmem_base + lane_off + TM_H + hb * HC;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                // 64 fp32 columns -> 32 packed bf16 columns, written over columns that have already been read
                uint32_t r[64], pk[32];
                tmem_ld_x32(ha + half * 64, r);
                tmem_ld_x32(ha + half * 64 + 32, r + 32);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 64; i += 4) {
                    const float4 bb = ld_shared_f32x4(sb1_u32 + (c * HC + half * 64 + i) * 4);
                    pk[(i >> 1)] = pack_bf16x2(fmaxf(__uint_as_float(r[i]) + bb.x, 0.f), fmaxf(__uint_as_float(r[i + 1]) + bb.y, 0.f));
                    pk[(i >> 1) + 1] = pack_bf16x2(fmaxf(__uint_as_float(r[i + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(r[i + 3]) + bb.w, 0.f));
                }
                tmem_st_x32(ha + half * 32, pk);
            }
            tmem_st_wait();
            tcgen05_fence_before();
            mbar_arrive(&p_ready[hb]);
            if (threadIdx.x == 6 * 32) PE_TR(2, g, 1);
        }
    } else {
        // ---------------- output epilogue: O + b2 (+ residuals) -> tokens, coalesced through a 2 KB slab ----------------
        const int quarter = warp & 3;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t sw = smem_u32(smem + OFF_SLAB) + (warp - 10) * 2048;
        const uint32_t sb2_u32 = smem_u32(sb2);
        for (int il = 0; il < my_tiles; ++il) {
            const int64_t tile0 = ((int64_t)blockIdx.x + (int64_t)il * gridDim.x) * TM;
            const int64_t t0w = tile0 + quarter * 32;
            const int r2_0 = p.residual2 ? (int)(t0w % p.res2_rows) : 0;
            // while the tile's MMAs run: pull this warp's residual rows (HBM) into L2, one 128-byte line per request
            if (p.residual) {
#pragma unroll 1
                for (int i = lane; i < 32 * 8; i += 32) {
                    const int64_t tt = t0w + (i >> 3);
                    if (tt < p.T) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + tt * EOUT + (i & 7) * 32));
                }
            }
            // Residual loads of one 32-column chunk in the coalesced (transposed) pattern: 16 INDEPENDENT, branch-free
            // requests issued back to back (rows past T are clamped, their results never stored).  A first version
            // added the two residuals inside a guarded loop; the loads then serialised on their ~800 clk latency and
            // this epilogue took 52 000 clk per tile, three times the tile's MMA time (tools/k1_trace.py).
            const int64_t t_last = p.T - 1;
            auto load_a = [&](int cc, float4* ra) {                    // `residual` (HBM, L2-prefetched above)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int rl = (q & 3) * 8 + (lane >> 2);
                    const int col = cc * 32 + (q >> 2) * 16 + (lane & 3) * 4;
                    const int64_t tt = (t0w + rl < t_last) ? t0w + rl : t_last;
                    ra[q] = p.residual ? __ldg(reinterpret_cast<const float4*>(p.residual + tt * EOUT + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            auto load_b = [&](int cc, float4* rb) {                    // `residual2` (broadcast rows, L2 resident)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int rl = (q & 3) * 8 + (lane >> 2);
                    const int col = cc * 32 + (q >> 2) * 16 + (lane & 3) * 4;
                    int r2 = r2_0 + rl;
                    r2 = (r2 >= p.res2_rows) ? r2 - p.res2_rows : r2;
                    rb[q] = __ldg(reinterpret_cast<const float4*>(p.residual2 + (int64_t)r2 * EOUT + col));
                }
            };
            float4 rnext[8];
            load_a(0, rnext);                                           // chunk 0's residual is requested before the tile is done
            mbar_wait(o_full, il & 1);
            tcgen05_fence_after();
            if (threadIdx.x == 10 * 32) PE_TR(3, il, 0);
#pragma unroll 1
            for (int cc = 0; cc < EOUT / 32; ++cc) {                    // 32-column chunk
                {
                    float4 res4[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) res4[q] = rnext[q];
                    uint32_t r[32];
                    tmem_ld_x32(tmem_base + lane_off + TM_O + cc * 32, r);
                    if (cc + 1 < EOUT / 32) load_a(cc + 1, rnext);      // next chunk's residual in flight under this chunk's work
                    if (p.residual2) {
                        float4 rb[8];
                        load_b(cc, rb);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            res4[q].x += rb[q].x; res4[q].y += rb[q].y; res4[q].z += rb[q].z; res4[q].w += rb[q].w;
                        }
                    }
                    tmem_ld_wait();
                    if (cc + 1 == EOUT / 32) {                          // accumulator fully read
                        tcgen05_fence_before(); mbar_arrive(o_free);
                        if (threadIdx.x == 10 * 32) PE_TR(3, il, 1);
                    }
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bz = ld_shared_f32x4(sb2_u32 + (cc * 32 + j) * 4);
                        v[j] = __uint_as_float(r[j]) + bz.x; v[j + 1] = __uint_as_float(r[j + 1]) + bz.y;
                        v[j + 2] = __uint_as_float(r[j + 2]) + bz.z; v[j + 3] = __uint_as_float(r[j + 3]) + bz.w;
                    }
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
                            const int64_t tt = t0w + rl;
                            if (tt >= p.T) continue;
                            const int col = cc * 32 + hh * 16 + piece * 4;
                            const float4 a = res4[hh * 4 + i];
                            x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
                            if (p.out) *reinterpret_cast<float4*>(p.out + tt * EOUT + col) = x;
                            if (p.out_bf16)
                                *reinterpret_cast<uint2*>(p.out_bf16 + tt * EOUT + col) =
                                    make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
                        }
                    }
                }
            }
            if (threadIdx.x == 10 * 32) PE_TR(3, il, 2);
        }
    }
    tcgen05_fence_before();
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
    p.trace = (long long*)tuning::debug_trace();
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
    const int64_t tiles = (p.T + TM - 1) / TM;
    PETR_REQUIRE(tiles < (1ll << 30), PETR_E_SHAPE, "petr_position_mlp_fused: too many tokens");
    p.num_tiles = (int)tiles;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    position_mlp_fused_kernel<<<grid, PE_THREADS, PE_SMEM, (cudaStream_t)stream>>>(tmW1, tmW2, p);
    return after_launch("petr_position_mlp_fused");
}
