// fp32 CUDA-core kernels of the PETR head hot path (sm_100a).
//
// These are the exact-arithmetic kernels: coordinate/mask generation (bit-exact against the CPU
// reference), the fp32 GEMM used where the 1e-5 / 2e-2 tolerances need fp32-class accuracy
// (query-side projections, branches, fp32 position-embedding path), LayerNorm, box decode and a
// flash-style fp32 attention (self-attention and the fp32 cross-attention path).  The Lk-side
// tensor-core work lives in tc_gemm.cu / tc_attention.cu.
#include "common.cuh"
#include <math.h>

namespace petr {

// ------------------------------------------------------------------------------------------
// a4: key-padding masks (petr_head.py:383-394)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
    // torch upsample_nearest: identity / >>1 shortcuts, else min(floorf(dst * (float)in/out), in-1)
    if (out_size == in_size) return dst;
    if (out_size == 2 * in_size) return dst >> 1;
    float scale = (float)in_size / (float)out_size;
    int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < in_size - 1 ? s : in_size - 1;
}

__global__ void key_padding_mask_kernel(const int32_t* __restrict__ img_hw, int BN, int pad_h,
                                        int pad_w, int H, int W, uint8_t* __restrict__ mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)BN * H * W;
    if (i >= total) return;
    int c = (int)(i % W);
    int r = (int)((i / W) % H);
    int bn = (int)(i / ((int64_t)W * H));
    int ih = img_hw[2 * bn], iw = img_hw[2 * bn + 1];
    int sr = nearest_src(r, pad_h, H), sc = nearest_src(c, pad_w, W);
    mask[i] = (sr < ih && sc < iw) ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// a1: frustum features (petr_head.py:286-331).  One warp per token (b,n,h,w); lane handles
// depths lane and lane+32 (D <= 64) or strides over D in general.
// ------------------------------------------------------------------------------------------
struct FrustumParams {
    int B, N, H, W, D, pad_h, pad_w, lid;
    float depth_start, bin_size;
    float lo[3], span[3];
};

__device__ __forceinline__ float inv_sigmoid_f(float x) {
    const float eps = 1e-5f;
    x = fminf(fmaxf(x, 0.f), 1.f);
    float x1 = fmaxf(x, eps);
    float x2 = fmaxf(__fsub_rn(1.f, x), eps);
    return logf(__fdiv_rn(x1, x2));
}

template <typename OutT>
__global__ void __launch_bounds__(256)
frustum_features_kernel(FrustumParams p, const float* __restrict__ img2lidar,
                        const uint8_t* __restrict__ masks, OutT* __restrict__ feat,
                        uint8_t* __restrict__ coords_mask, float* __restrict__ coords3d) {
    __shared__ float sM[8][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t T = (int64_t)p.B * p.N * p.H * p.W;
    const int64_t t = (int64_t)blockIdx.x * 8 + warp;
    const bool live = t < T;
    int w = 0, h = 0, bn = 0;
    if (live) {
        w = (int)(t % p.W);
        h = (int)((t / p.W) % p.H);
        bn = (int)(t / ((int64_t)p.W * p.H));
        if (lane < 12) sM[warp][lane] = img2lidar[(int64_t)bn * 16 + lane];   // rows 0..2
    }
    __syncwarp();
    if (!live) return;
    const float* M = sM[warp];
    // pixel-corner grids: arange().float() * pad / size  (petr_head.py:290-291)
    const float u = __fdiv_rn(__fmul_rn((float)w, (float)p.pad_w), (float)p.W);
    const float v = __fdiv_rn(__fmul_rn((float)h, (float)p.pad_h), (float)p.H);
    int bad = 0;
    for (int d = lane; d < p.D; d += 32) {
        float fd = (float)d, depth;
        if (p.lid)   // depth_start + bin_size * index * (index + 1)  (petr_head.py:294-297)
            depth = __fadd_rn(p.depth_start, __fmul_rn(__fmul_rn(p.bin_size, fd), __fadd_rn(fd, 1.f)));
        else         // depth_start + bin_size * index               (petr_head.py:299-302)
            depth = __fadd_rn(p.depth_start, __fmul_rn(p.bin_size, fd));
        const float dm = fmaxf(depth, 1e-5f);
        const float p0 = __fmul_rn(u, dm), p1 = __fmul_rn(v, dm), p2 = depth;
        float c[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            // ((m0*p0 + m1*p1) + m2*p2) + m3*1 with separately rounded ops (petr_head.py:319)
            float acc = __fmul_rn(M[4 * i + 0], p0);
            acc = __fadd_rn(acc, __fmul_rn(M[4 * i + 1], p1));
            acc = __fadd_rn(acc, __fmul_rn(M[4 * i + 2], p2));
            acc = __fadd_rn(acc, M[4 * i + 3]);
            c[i] = __fdiv_rn(__fsub_rn(acc, p.lo[i]), p.span[i]);          // :320-325
            bad += (c[i] > 1.0f) | (c[i] < 0.0f);                          // :327
        }
        if (coords3d) {
            float* o = coords3d + (t * p.D + d) * 3;
            o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
        }
        OutT* f = feat + t * (int64_t)(3 * p.D) + 3 * d;
        if constexpr (sizeof(OutT) == 4) {
            f[0] = inv_sigmoid_f(c[0]); f[1] = inv_sigmoid_f(c[1]); f[2] = inv_sigmoid_f(c[2]);
        } else {
            f[0] = __float2bfloat16_rn(inv_sigmoid_f(c[0]));
            f[1] = __float2bfloat16_rn(inv_sigmoid_f(c[1]));
            f[2] = __float2bfloat16_rn(inv_sigmoid_f(c[2]));
        }
    }
    if (coords_mask) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
        if (lane == 0) {
            // sum > D * 0.5 (python float compare, petr_head.py:328)
            uint8_t m = ((float)bad > (float)p.D * 0.5f) ? 1 : 0;
            if (masks) m |= (masks[t] != 0);
            coords_mask[t] = m;
        }
    }
}

// ------------------------------------------------------------------------------------------
// a3: SinePositionalEncoding3D features (positional_encoding.py:58-100).  One block per
// (b, n, h) feature row.
// ------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256)
sine3d_features_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ dim_t,
                       int B, int N, int H, int W, int F, int normalize, float scale, float eps,
                       float offset, OutT* __restrict__ out) {
    extern __shared__ float sm[];            // [3][W] values, then [F] dim_t
    float* vn = sm;
    float* vy = sm + W;
    float* vx = sm + 2 * W;
    float* sdt = sm + 3 * W;
    const int h = blockIdx.x % H;
    const int n = (blockIdx.x / H) % N;
    const int b = blockIdx.x / (H * N);
    const int64_t HW = (int64_t)H * W;
    const uint8_t* mb = mask + (int64_t)b * N * HW;
    for (int i = threadIdx.x; i < F; i += blockDim.x) sdt[i] = dim_t[i];
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        int en = 0, ln = 0, ey = 0, ly = 0;
        for (int k = 0; k < N; ++k) {
            int nm = mb[k * HW + (int64_t)h * W + w] ? 0 : 1;
            ln += nm;
            if (k <= n) en += nm;
        }
        for (int k = 0; k < H; ++k) {
            int nm = mb[n * HW + (int64_t)k * W + w] ? 0 : 1;
            ly += nm;
            if (k <= h) ey += nm;
        }
        float fn = (float)en, fy = (float)ey;
        if (normalize) {
            fn = __fmul_rn(__fdiv_rn(__fadd_rn(fn, offset), __fadd_rn((float)ln, eps)), scale);
            fy = __fmul_rn(__fdiv_rn(__fadd_rn(fy, offset), __fadd_rn((float)ly, eps)), scale);
        }
        vn[w] = fn;
        vy[w] = fy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        const uint8_t* row = mb + n * HW + (int64_t)h * W;
        for (int w = 0; w < W; ++w) { acc += row[w] ? 0 : 1; vx[w] = (float)acc; }
        if (normalize) {
            float last = __fadd_rn((float)acc, eps);
            for (int w = 0; w < W; ++w)
                vx[w] = __fmul_rn(__fdiv_rn(__fadd_rn(vx[w], offset), last), scale);
        }
    }
    __syncthreads();
    const int C = 3 * F, half = F / 2;
    OutT* orow = out + ((int64_t)blockIdx.x * W) * C;
    for (int i = threadIdx.x; i < W * C; i += blockDim.x) {
        int w = i / C, ch = i % C;
        int axis = ch / F, j = ch % F;
        float val = axis == 0 ? vn[w] : (axis == 1 ? vy[w] : vx[w]);
        float r;
        if constexpr (sizeof(OutT) == 4) {
            if (j < half) r = sinf(__fdiv_rn(val, sdt[2 * j]));
            else          r = cosf(__fdiv_rn(val, sdt[2 * (j - half) + 1]));
            orow[i] = r;
        } else {
            // bf16 output (8 mantissa bits): the fast intrinsics (abs error ~1e-6 on [0, 2*pi]) are exact enough
            if (j < half) r = __sinf(__fdiv_rn(val, sdt[2 * j]));
            else          r = __cosf(__fdiv_rn(val, sdt[2 * (j - half) + 1]));
            orow[i] = __float2bfloat16_rn(r);
        }
    }
}

// a6: pos2posemb3d (petr_head.py:31-43)
__global__ void pos2posemb3d_kernel(const float* __restrict__ ref, const float* __restrict__ dim_t,
                                    int Q, int F, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int C = 3 * F;
    if (i >= Q * C) return;
    int q = i / C, ch = i % C, blk = ch / F, j = ch % F;
    const int axis_of_blk[3] = {1, 0, 2};                       // cat((pos_y, pos_x, pos_z))
    float p = __fmul_rn(ref[q * 3 + axis_of_blk[blk]], 6.283185307179586f);
    float a = __fdiv_rn(p, dim_t[j]);
    out[i] = (j & 1) ? cosf(a) : sinf(a);
}

// ------------------------------------------------------------------------------------------
// fp32 GEMM: C = epi((A + A2) W^T + bias) + residual.  128x64 tile, BK 16, 256 threads, 8x4/thread
// ------------------------------------------------------------------------------------------
constexpr int GBM = 128, GBN = 64, GBK = 16;

struct GemmParams {
    const float* A; int lda; int a_layout; int hw;
    const float* A2; int a2_rows;
    const float* W; const float* bias; const float* residual; const float* gate;
    int act; float* C; __nv_bfloat16* Cb; int ldc; int M, N, K;
};

__device__ __forceinline__ float gemm_load_a(const GemmParams& p, int m, int k) {
    if (m >= p.M || k >= p.K) return 0.f;
    float v;
    if (p.a_layout == PETR_A_NCHW) {
        int g = m / p.hw, r = m - g * p.hw;
        v = p.A[((int64_t)g * p.K + k) * p.hw + r];
    } else {
        v = p.A[(int64_t)m * p.lda + k];
    }
    if (p.A2) v += p.A2[(int64_t)(m % p.a2_rows) * p.lda + k];
    return v;
}

__global__ void __launch_bounds__(256)
gemm_f32_kernel(GemmParams p) {
    __shared__ __align__(16) float As[2][GBK][GBM + 4];
    __shared__ __align__(16) float Bs[2][GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int tm = (tid / 16) * 8, tn = (tid % 16) * 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const bool vecA = (p.a_layout == PETR_A_ROWMAJOR) && ((p.lda & 3) == 0) && ((p.K & 3) == 0) &&
                      aligned16(p.A) && (!p.A2 || aligned16(p.A2));
    const bool vecB = ((p.K & 3) == 0) && aligned16(p.W);
    float ra[8], rb[4];

    auto load_tiles = [&](int k0) {
        if (p.a_layout == PETR_A_NCHW) {
            // contiguous along m for fixed k: thread -> (k = tid/16 .. , m = tid%16*8..)
            int k = tid / 16, mm = (tid % 16) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) ra[i] = gemm_load_a(p, m0 + mm + i, k0 + k);
        } else if (vecA) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                int idx = tid + r * 256;                 // 512 float4: 128 rows x 4
                int m = m0 + idx / 4, k = k0 + (idx % 4) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < p.M && k < p.K) {
                    v = *reinterpret_cast<const float4*>(p.A + (int64_t)m * p.lda + k);
                    if (p.A2) {
                        float4 w = *reinterpret_cast<const float4*>(p.A2 + (int64_t)(m % p.a2_rows) * p.lda + k);
                        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                    }
                }
                ra[r * 4 + 0] = v.x; ra[r * 4 + 1] = v.y; ra[r * 4 + 2] = v.z; ra[r * 4 + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                int idx = tid + r * 256;                 // 2048 scalars: 128 rows x 16
                ra[r] = gemm_load_a(p, m0 + idx / 16, k0 + idx % 16);
            }
        }
        if (vecB) {
            int n = n0 + tid / 4, k = k0 + (tid % 4) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < p.N && k < p.K) v = *reinterpret_cast<const float4*>(p.W + (int64_t)n * p.K + k);
            rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + r * 256;
                int n = n0 + idx / 16, k = k0 + idx % 16;
                rb[r] = (n < p.N && k < p.K) ? p.W[(int64_t)n * p.K + k] : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
        if (p.a_layout == PETR_A_NCHW) {
            int k = tid / 16, mm = (tid % 16) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) As[buf][k][mm + i] = ra[i];
        } else if (vecA) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                int idx = tid + r * 256;
                int m = idx / 4, k = (idx % 4) * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) As[buf][k + j][m] = ra[r * 4 + j];
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                int idx = tid + r * 256;
                As[buf][idx % 16][idx / 16] = ra[r];
            }
        }
        if (vecB) {
            int n = tid / 4, k = (tid % 4) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[buf][k + j][n] = rb[j];
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int idx = tid + r * 256;
                Bs[buf][idx % 16][idx / 16] = rb[r];
            }
        }
    };

    const int nk = (p.K + GBK - 1) / GBK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tiles((kb + 1) * GBK);
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tn]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + tm + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tn + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (p.bias) v += p.bias[n];
            int64_t o = (int64_t)m * p.ldc + n;
            if (p.act == PETR_ACT_RELU) v = fmaxf(v, 0.f);
            else if (p.act == PETR_ACT_SIGMOID_GATE) v = p.gate[o] * (1.f / (1.f + expf(-v)));
            if (p.residual) v += p.residual[o];
            if (p.C) p.C[o] = v;
            if (p.Cb) p.Cb[o] = __float2bfloat16_rn(v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// fp32 flash-style attention, head_dim 32 (the exact-arithmetic parity path and the training forward).
// One thread owns one query row (q, the online-softmax state m / l and o[32] in registers); a block of 64 rows streams
// 32-key tiles of K and V from shared memory -- every lane reads the SAME key row, so the loads are broadcasts -- in
// groups of 8 keys (one running-maximum update and one rescale of o per group).  All dot products / axpys are packed
// fma.rn.f32x2 (two fp32 FMAs per issued instruction on sm_100: the FMA pipe stays at 128 lanes/clk/SM but the issue
// slots the shared-memory loads need are freed; tools/probes/probe_ffma2.cu).  1.6x the quad-per-row version it
// replaces (shuffle merges, scalar FMAs): 13.5 -> 8.4 ms for the six cross-attentions of a 1600x640 sample.
// ------------------------------------------------------------------------------------------
constexpr int ATT_ROWS = 64, ATT_KT = 32;

__global__ void __launch_bounds__(ATT_ROWS)
attention_f32_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
                     const float* __restrict__ V, int ldv, const uint8_t* __restrict__ key_mask,
                     float* __restrict__ O, int ldo, __nv_bfloat16* __restrict__ O_split, int E,
                     int Lq, int Lk, float scale_log2, const float* __restrict__ attn_bias,
                     int64_t bias_bh_stride, float* __restrict__ lse, float drop_p, uint64_t seed,
                     float* __restrict__ partial, int ksplit) {
    __shared__ __align__(16) float sKb[2][ATT_KT][32];
    __shared__ __align__(16) float sVb[2][ATT_KT][32];
    __shared__ float sBiasb[2][ATT_KT];
    // ksplit > 1: the key axis is split over `ksplit` blocks per (row block, head, sample) -- 900 query rows x 8 heads
    // are 113 warps per sample, too few for 148 SMs -- and every part writes its un-normalised (o, m, l) to `partial`
    // in the layout petr_attention_merge reads
    const int head = blockIdx.y, b = blockIdx.z / ksplit, part = blockIdx.z - b * ksplit;
    const int q = blockIdx.x * ATT_ROWS + threadIdx.x;
    const bool live = q < Lq;
    const int bh = b * gridDim.y + head;
    uint64_t q2[16], o2[16];
    {
        const float* qp = Q + ((int64_t)b * Lq + (live ? q : 0)) * ldq + head * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(qp + i);
            q2[i >> 1] = pack2(t.x * scale_log2, t.y * scale_log2);
            q2[(i >> 1) + 1] = pack2(t.z * scale_log2, t.w * scale_log2);
            o2[i >> 1] = o2[(i >> 1) + 1] = 0ull;
        }
    }
    float m = -INFINITY, l = 0.f;
    // attn_mask of torch.nn.MultiheadAttention as an additive fp32 bias [Lq, Lk] (bias_bh_stride 0) or
    // [B*heads, Lq, Lk]; attention-probability dropout (training) is a counter-based hash of (row, key)
    const float* brow = attn_bias ? attn_bias + (int64_t)bh * bias_bh_stride + (int64_t)(live ? q : 0) * Lk : nullptr;
    const uint64_t drop_row = ((uint64_t)bh * Lq + (live ? q : 0)) * (uint64_t)Lk;
    const float drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const float* Kb = K + (int64_t)b * Lk * ldk + head * 32;
    const float* Vb = V + (int64_t)b * Lk * ldv + head * 32;
    const int tiles = (Lk + ATT_KT - 1) / ATT_KT;
    const int k_begin = (int)((int64_t)tiles * part / ksplit) * ATT_KT;
    const int k_end = min(Lk, (int)((int64_t)tiles * (part + 1) / ksplit) * ATT_KT);
    // the next key tile streams into the other buffer (cp.async) while this one is processed
    auto issue_tile = [&](int k0, int buf) {
        for (int i = threadIdx.x; i < ATT_KT * 8; i += ATT_ROWS) {
            const int r = i >> 3, c4 = (i & 7) * 4;
            const bool valid = k0 + r < k_end;
            const int64_t row = valid ? k0 + r : 0;
            cp_async16(&sKb[buf][r][c4], Kb + row * ldk + c4, valid);
            cp_async16(&sVb[buf][r][c4], Vb + row * ldv + c4, valid);
        }
        if (threadIdx.x < ATT_KT) {
            const int r = threadIdx.x;
            const bool dead = (k0 + r >= k_end) || (key_mask && key_mask[(int64_t)b * Lk + k0 + r]);
            sBiasb[buf][r] = dead ? -INFINITY : 0.f;
        }
        cp_async_commit();
    };
    if (k_begin < k_end) issue_tile(k_begin, 0);
    for (int k0 = k_begin, it = 0; k0 < k_end; k0 += ATT_KT, ++it) {
        const int buf = it & 1;
        if (k0 + ATT_KT < k_end) { issue_tile(k0 + ATT_KT, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();                               // tile has landed for every thread
        const float (*sK)[32] = sKb[buf];
        const float (*sV)[32] = sVb[buf];
        const float* sBias = sBiasb[buf];
#pragma unroll 1
        for (int g = 0; g < ATT_KT / 8; ++g) {
            float sc[8];
            float mx = m;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int r = g * 8 + jj;
                uint64_t a0 = 0ull, a1 = 0ull;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const ulonglong2 ka = *reinterpret_cast<const ulonglong2*>(&sK[r][2 * i]);
                    const ulonglong2 kb = *reinterpret_cast<const ulonglong2*>(&sK[r][2 * i + 4]);
                    a0 = fma2(q2[i], ka.x, a0); a1 = fma2(q2[i + 1], ka.y, a1);
                    a0 = fma2(q2[i + 2], kb.x, a0); a1 = fma2(q2[i + 3], kb.y, a1);
                }
                const float2 u = unpack2(a0), w = unpack2(a1);
                float sj = ((u.x + u.y) + (w.x + w.y)) + sBias[r];
                if (brow && k0 + r < Lk) sj += brow[k0 + r] * 1.4426950408889634f;
                sc[jj] = sj;
                mx = fmaxf(mx, sj);
            }
            if (mx == -INFINITY) continue;             // everything so far masked
            if (mx > m) {                              // new running maximum: rescale the state once per group
                const float alpha = exp2f(m - mx);     // m = -inf -> 0
                const uint64_t al2 = pack2(alpha, alpha);
                l *= alpha;
#pragma unroll
                for (int i = 0; i < 16; ++i) o2[i] = mul2(o2[i], al2);
                m = mx;
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int r = g * 8 + jj;
                float pj = exp2f(sc[jj] - m);
                l += pj;
                if (drop_p > 0.f) pj = dropout_keep(seed, drop_row + (uint64_t)(k0 + r), drop_p) ? pj * drop_scale : 0.f;
                const uint64_t p2 = pack2(pj, pj);
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const ulonglong2 vv = *reinterpret_cast<const ulonglong2*>(&sV[r][2 * i]);
                    o2[i] = fma2(p2, vv.x, o2[i]);
                    o2[i + 1] = fma2(p2, vv.y, o2[i + 1]);
                }
            }
        }
        __syncthreads();                               // buffer free before the tile after next is issued into it
    }
    if (!live) return;
    if (ksplit > 1) {
        const int64_t rows = (int64_t)gridDim.z / ksplit * gridDim.y * Lq;
        const int64_t r = (int64_t)bh * Lq + q;
        float* po = partial + (int64_t)part * rows * 34 + r * 32;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const float2 t0 = unpack2(o2[i]), t1 = unpack2(o2[i + 1]);
            *reinterpret_cast<float4*>(po + 2 * i) = make_float4(t0.x, t0.y, t1.x, t1.y);
        }
        float* pml = partial + (int64_t)part * rows * 34 + rows * 32 + r * 2;
        pml[0] = m;
        pml[1] = l;
        return;
    }
    if (lse) lse[(int64_t)bh * Lq + q] = m + log2f(l);         // log2-domain log-sum-exp of the scaled scores (backward)
    const float inv = 1.f / l;                                 // l == 0 (all keys masked) -> NaN, as torch
    float r32[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float2 t = unpack2(o2[i]); r32[2 * i] = t.x * inv; r32[2 * i + 1] = t.y * inv; }
    if (O) {
        float* op = O + ((int64_t)b * Lq + q) * ldo + head * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(r32[i], r32[i + 1], r32[i + 2], r32[i + 3]);
    }
    if (O_split) {                                             // bf16 [hi | lo] rows of width 2E (bf16x3 GEMM input)
        __nv_bfloat16* sp = O_split + ((int64_t)b * Lq + q) * (2 * E) + head * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(r32[i]);
            sp[i] = hi;
            sp[E + i] = __float2bfloat16_rn(r32[i] - __bfloat162float(hi));
        }
    }
}

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, E multiple of 32, E <= 2048
// ------------------------------------------------------------------------------------------
template <int PER>
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ X, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ Y, int rows,
                     float eps, int relu, int nan_to_num) {
    constexpr int E = PER * 32;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* x = X + (int64_t)row * E;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = x[lane + 32 * i]; s += v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)E;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)E + eps);
    float* y = Y + (int64_t)row * E;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int c = lane + 32 * i;
        float r = (v[i] - mean) * rstd * gamma[c] + beta[c];
        if (relu) r = fmaxf(r, 0.f);       // only used on finite branch inputs
        if (nan_to_num) {
            if (isnan(r)) r = 0.f;
            else if (isinf(r)) r = r > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
        }
        y[c] = r;
    }
}

// Fused LayerNorm for the decoder residual stream (E = 256, one warp per row): optional outputs
//   Y fp32, split(Y) and split(Y + qpos[row % qpos_rows]) as bf16 [hi | lo] rows (inputs of the
//   bf16x3 tensor-core GEMMs), and a second LayerNorm Z = LN2(Y) (the decoder's post_norm on a
//   copy, petr_transformer.py:444) with nan_to_num (petr_head.py:435) as fp32 and split bf16.
struct LnMultiParams {
    const float* X; const float* gamma; const float* beta; float eps;
    const float* qpos; int qpos_rows;
    float* Y; __nv_bfloat16* Y_split; __nv_bfloat16* Ypos_split;
    const float* gamma2; const float* beta2; float* Z; __nv_bfloat16* Z_split;
    int rows; int relu;
    int nparts; int64_t part_stride;     // X = sum of nparts arrays (split-K partial products of the preceding GEMM)
};

__device__ __forceinline__ void store_split_elem(__nv_bfloat16* base, int E, int c, float v) {
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    base[c] = hi;
    base[E + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// 8 consecutive fp32 values -> bf16 hi / residual rows (one 16-byte store each).  Paired converts
// (F2FP, ALU pipe) instead of scalar ones (F2F, quarter-rate XU pipe).
__device__ __forceinline__ void store_split8(__nv_bfloat16* hp, __nv_bfloat16* lp, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 hv = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        h[i] = *reinterpret_cast<uint32_t*>(&hv);
        __nv_bfloat162 lv = __floats2bfloat162_rn(v[2 * i] - __uint_as_float(h[i] << 16),
                                                  v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u));
        l[i] = *reinterpret_cast<uint32_t*>(&lv);
    }
    *reinterpret_cast<uint4*>(hp) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lp) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void load8(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// one warp per row of E = 256; lane owns columns [8 lane, 8 lane + 8): 16-byte loads and stores
__global__ void __launch_bounds__(256)
layernorm_multi_kernel(LnMultiParams p) {
    constexpr int PER = 8, E = 256;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    if (row >= p.rows) return;
    const int c0 = lane * PER;
    float v[PER], y[PER], g[PER], bt[PER];
    load8(p.X + (int64_t)row * E + c0, v);
    for (int pt = 1; pt < p.nparts; ++pt) {            // fixed order: deterministic
        float w[PER];
        load8(p.X + pt * p.part_stride + (int64_t)row * E + c0, w);
#pragma unroll
        for (int i = 0; i < PER; ++i) v[i] += w[i];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float mean = s / (float)E, q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    float rstd = rsqrtf(q / (float)E + p.eps);
    load8(p.gamma + c0, g);
    load8(p.beta + c0, bt);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        y[i] = (v[i] - mean) * rstd * g[i] + bt[i];
        if (p.relu) y[i] = fmaxf(y[i], 0.f);
    }
    if (p.Y) store8(p.Y + (int64_t)row * E + c0, y);
    if (p.Y_split) {
        __nv_bfloat16* sp = p.Y_split + (int64_t)row * 2 * E + c0;
        store_split8(sp, sp + E, y);
    }
    if (p.Ypos_split) {
        float qp[PER];
        load8(p.qpos + (int64_t)(row % p.qpos_rows) * E + c0, qp);
#pragma unroll
        for (int i = 0; i < PER; ++i) qp[i] += y[i];
        __nv_bfloat16* sp = p.Ypos_split + (int64_t)row * 2 * E + c0;
        store_split8(sp, sp + E, qp);
    }
    if (p.gamma2) {
        s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) s += y[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        mean = s / (float)E; q = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { float d = y[i] - mean; q += d * d; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        rstd = rsqrtf(q / (float)E + p.eps);
        load8(p.gamma2 + c0, g);
        load8(p.beta2 + c0, bt);
        float r[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            r[i] = (y[i] - mean) * rstd * g[i] + bt[i];
            if (isnan(r[i])) r[i] = 0.f;
            else if (isinf(r[i])) r[i] = r[i] > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
        }
        if (p.Z) store8(p.Z + (int64_t)row * E + c0, r);
        if (p.Z_split) {
            __nv_bfloat16* sp = p.Z_split + (int64_t)row * 2 * E + c0;
            store_split8(sp, sp + E, r);
        }
    }
}

// fp32 rows -> bf16 [hi | lo] rows, optionally adding a broadcast row table first, with optional
// ReLU (E arbitrary): split(relu(X + add[row % add_rows]))
__global__ void split_rows_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ add,
                                  int add_rows, __nv_bfloat16* __restrict__ out, int64_t rows, int E, int relu) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * E) return;
    int64_t r = i / E;
    int c = (int)(i - r * E);
    float v = X[r * ldx + c];
    if (add) v += add[(r % add_rows) * (int64_t)E + c];
    if (relu) v = fmaxf(v, 0.f);
    store_split_elem(out + r * 2 * E, E, c, v);
}

// the same, eight columns per thread (E % 8 == 0, ldx % 4 == 0, 16-byte aligned pointers)
__global__ void split_rows8_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ add,
                                   int add_rows, __nv_bfloat16* __restrict__ out, int64_t rows, int E, int relu) {
    const int E8 = E >> 3;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * E8) return;
    int64_t r = i / E8;
    int c = (int)(i - r * E8) * 8;
    float v[8];
    load8(X + r * ldx + c, v);
    if (add) {
        float a[8];
        load8(add + (r % add_rows) * (int64_t)E + c, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    store_split8(out + r * 2 * E + c, out + r * 2 * E + E + c, v);
}

// fp32 weight (N,K) -> the W operand of a bf16x3 GEMM: rows [hi | hi | lo] of W (transpose 0: (N, 3K)) or of its
// transpose (transpose 1: (K, 3N), the operand of the input-gradient GEMM dA = dY W).
__global__ void x3_weight_kernel(const float* __restrict__ W, int N, int K, int transpose,
                                 __nv_bfloat16* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int C = transpose ? N : K;
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float v = transpose ? W[(int64_t)c * K + r] : W[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16* o = out + r * 3 * C;
    o[c] = hi;
    o[C + c] = hi;
    o[2 * C + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// out_bf16[r, :] = bf16(X[r, :] + add[r % add_rows, :]); E % 4 == 0 (key-side K input: tokens +
// the batch-invariant sine embedding)
__global__ void add_rows_bf16_kernel(const float* __restrict__ X, const float* __restrict__ add, int add_rows,
                                     __nv_bfloat16* __restrict__ out, int64_t rows, int E4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * E4) return;
    int64_t r = i / E4;
    int c = (int)(i - r * E4);
    float4 v = reinterpret_cast<const float4*>(X)[i];
    const float4 a = __ldg(reinterpret_cast<const float4*>(add) + (r % add_rows) * E4 + c);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x + a.x, v.y + a.y), hi = __floats2bfloat162_rn(v.z + a.z, v.w + a.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = o;
}

// Small host->device staging through the SMs: `src` is pinned (UVA-mapped) host memory.  The few
// hundred bytes of per-call meta data (img2lidar, img_shape, timestamps) must not queue on the H2D
// copy engine behind a 100s-of-MB feature upload that runs concurrently on another stream.
__global__ void stage_small_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int nwords) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// a12: reference add + sigmoid + pc_range affine (petr_head.py:440-460)
struct DecodeParams { float pc[6]; };

__global__ void decode_boxes_kernel(float* __restrict__ reg, const float* __restrict__ ref,
                                    DecodeParams dp, const float* __restrict__ time_div,
                                    int64_t rows, int B, int Q, int cs) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int q = (int)(r % Q);
    int b = (int)((r / Q) % B);
    float* t = reg + r * cs;
    float rx = inv_sigmoid_f(ref[q * 3 + 0]), ry = inv_sigmoid_f(ref[q * 3 + 1]),
          rz = inv_sigmoid_f(ref[q * 3 + 2]);
    float x = 1.f / (1.f + expf(-(t[0] + rx)));
    float y = 1.f / (1.f + expf(-(t[1] + ry)));
    float z = 1.f / (1.f + expf(-(t[4] + rz)));
    t[0] = x * (dp.pc[3] - dp.pc[0]) + dp.pc[0];
    t[1] = y * (dp.pc[4] - dp.pc[1]) + dp.pc[1];
    t[4] = z * (dp.pc[5] - dp.pc[2]) + dp.pc[2];
    if (time_div) {
        float d = time_div[b];
        for (int c = 8; c < cs; ++c) t[c] = t[c] / d;
    }
}

}  // namespace petr

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace petr;

extern "C" int petr_key_padding_mask(const int32_t* img_hw, int B, int N, int pad_h, int pad_w,
                                     int H, int W, uint8_t* mask, void* stream) {
    PETR_REQUIRE(img_hw && mask, PETR_E_BADARG, "petr_key_padding_mask: null pointer");
    PETR_REQUIRE(B > 0 && N > 0 && H > 0 && W > 0 && pad_h > 0 && pad_w > 0, PETR_E_BADARG,
                 "petr_key_padding_mask: non-positive dimension");
    int64_t total = (int64_t)B * N * H * W;
    key_padding_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        img_hw, B * N, pad_h, pad_w, H, W, mask);
    return after_launch("petr_key_padding_mask");
}

extern "C" int petr_frustum_features(const float* img2lidar, const uint8_t* masks, int B, int N,
                                     int H, int W, int D, int pad_h, int pad_w, int lid,
                                     float depth_start, float bin_size, const float* h_range_lo,
                                     const float* h_range_span, void* feat, int feat_dtype,
                                     uint8_t* coords_mask, float* coords3d, void* stream) {
    PETR_REQUIRE(img2lidar && feat && h_range_lo && h_range_span, PETR_E_BADARG,
                 "petr_frustum_features: null pointer");
    PETR_REQUIRE(B > 0 && N > 0 && H > 0 && W > 0 && D > 0, PETR_E_BADARG,
                 "petr_frustum_features: non-positive dimension");
    PETR_REQUIRE(feat_dtype == PETR_DTYPE_F32 || feat_dtype == PETR_DTYPE_BF16, PETR_E_BADARG,
                 "petr_frustum_features: bad dtype %d", feat_dtype);
    FrustumParams p;
    p.B = B; p.N = N; p.H = H; p.W = W; p.D = D; p.pad_h = pad_h; p.pad_w = pad_w; p.lid = lid;
    p.depth_start = depth_start; p.bin_size = bin_size;
    for (int i = 0; i < 3; ++i) { p.lo[i] = h_range_lo[i]; p.span[i] = h_range_span[i]; }
    int64_t T = (int64_t)B * N * H * W;
    unsigned grid = (unsigned)((T + 7) / 8);
    if (feat_dtype == PETR_DTYPE_F32)
        frustum_features_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(
            p, img2lidar, masks, (float*)feat, coords_mask, coords3d);
    else
        frustum_features_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
            p, img2lidar, masks, (__nv_bfloat16*)feat, coords_mask, coords3d);
    return after_launch("petr_frustum_features");
}

extern "C" int petr_sine3d_features(const uint8_t* mask, const float* dim_t, int B, int N, int H,
                                    int W, int num_feats, int normalize, float scale, float eps,
                                    float offset, void* out, int out_dtype, void* stream) {
    PETR_REQUIRE(mask && dim_t && out, PETR_E_BADARG, "petr_sine3d_features: null pointer");
    PETR_REQUIRE(B > 0 && N > 0 && H > 0 && W > 0 && num_feats > 0 && (num_feats % 2) == 0,
                 PETR_E_BADARG, "petr_sine3d_features: bad dimension");
    size_t smem = (size_t)(3 * W + num_feats) * sizeof(float);
    PETR_REQUIRE(smem <= 48 * 1024, PETR_E_SHAPE, "petr_sine3d_features: W too large");
    unsigned grid = (unsigned)(B * N * H);
    if (out_dtype == PETR_DTYPE_F32)
        sine3d_features_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(
            mask, dim_t, B, N, H, W, num_feats, normalize, scale, eps, offset, (float*)out);
    else if (out_dtype == PETR_DTYPE_BF16)
        sine3d_features_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>(
            mask, dim_t, B, N, H, W, num_feats, normalize, scale, eps, offset, (__nv_bfloat16*)out);
    else
        return fail(PETR_E_BADARG, "petr_sine3d_features: bad dtype %d", out_dtype);
    return after_launch("petr_sine3d_features");
}

extern "C" int petr_pos2posemb3d(const float* ref, const float* dim_t, int Q, int num_feats,
                                 float* out, void* stream) {
    PETR_REQUIRE(ref && dim_t && out, PETR_E_BADARG, "petr_pos2posemb3d: null pointer");
    PETR_REQUIRE(Q > 0 && num_feats > 0, PETR_E_BADARG, "petr_pos2posemb3d: bad dimension");
    int total = Q * 3 * num_feats;
    pos2posemb3d_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ref, dim_t, Q, num_feats, out);
    return after_launch("petr_pos2posemb3d");
}

// Skinny-N fp32 GEMM for the last layer of the branches (256 -> 10 over
// 6 * B * 900 rows, petr_head.py:226-247): the tiled kernel above spends 110 us per launch at B = 16
// on an 88 MB read (0.8 TB/s).  Here a warp keeps its slice of W in registers (lane l: columns
// [8 l, 8 l + 8) of all N <= 10 rows of W), streams rows of A with two 16-byte loads per lane, and
// reduces the N partial sums with shuffles; one row per warp iteration, 8 warps per CTA.
constexpr int SKINNY_MAX_N = 10;
__global__ void __launch_bounds__(256)
gemm_f32_skinny_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                       const float* __restrict__ bias, const float* __restrict__ residual,
                       float* __restrict__ C, int ldc, int M, int N, int relu) {
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * 8;
    float w[SKINNY_MAX_N][8];
#pragma unroll
    for (int n = 0; n < SKINNY_MAX_N; ++n) {
        if (n < N) {
            const float4 a = *reinterpret_cast<const float4*>(W + (int64_t)n * 256 + lane * 8);
            const float4 b = *reinterpret_cast<const float4*>(W + (int64_t)n * 256 + lane * 8 + 4);
            w[n][0] = a.x; w[n][1] = a.y; w[n][2] = a.z; w[n][3] = a.w;
            w[n][4] = b.x; w[n][5] = b.y; w[n][6] = b.z; w[n][7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[n][i] = 0.f;
        }
    }
    const float my_bias = (bias && lane < N) ? bias[lane] : 0.f;
    for (int m = warp_global; m < M; m += warps_total) {
        const float* ap = A + (int64_t)m * lda + lane * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(ap), a1 = *reinterpret_cast<const float4*>(ap + 4);
        float acc[SKINNY_MAX_N];
#pragma unroll
        for (int n = 0; n < SKINNY_MAX_N; ++n) {
            float t = a0.x * w[n][0];
            t = fmaf(a0.y, w[n][1], t); t = fmaf(a0.z, w[n][2], t); t = fmaf(a0.w, w[n][3], t);
            t = fmaf(a1.x, w[n][4], t); t = fmaf(a1.y, w[n][5], t); t = fmaf(a1.z, w[n][6], t);
            acc[n] = fmaf(a1.w, w[n][7], t);
        }
        // butterfly reduction; afterwards every lane holds all N sums
#pragma unroll
        for (int n = 0; n < SKINNY_MAX_N; ++n) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
        }
        float mine = 0.f;                                    // lane n keeps sum n
#pragma unroll
        for (int n = 0; n < SKINNY_MAX_N; ++n) mine = (lane == n) ? acc[n] : mine;
        if (lane < N) {
            float v = mine + my_bias;
            if (relu) v = fmaxf(v, 0.f);
            if (residual) v += residual[(int64_t)m * ldc + lane];
            C[(int64_t)m * ldc + lane] = v;
        }
    }
}

extern "C" int petr_gemm_f32(const float* A, int lda, int a_layout, int hw, const float* A2,
                             int a2_rows, const float* W, const float* bias, const float* residual,
                             const float* gate, int act, float* C, void* C_bf16, int ldc, int M,
                             int N, int K, void* stream) {
    PETR_REQUIRE(A && W && (C || C_bf16), PETR_E_BADARG, "petr_gemm_f32: null pointer");
    PETR_REQUIRE(M > 0 && N > 0 && K > 0 && ldc >= N, PETR_E_BADARG, "petr_gemm_f32: bad dimension");
    PETR_REQUIRE(a_layout == PETR_A_ROWMAJOR || a_layout == PETR_A_NCHW, PETR_E_BADARG,
                 "petr_gemm_f32: bad a_layout");
    if (a_layout == PETR_A_NCHW)
        PETR_REQUIRE(hw > 0 && M % hw == 0 && !A2, PETR_E_SHAPE, "petr_gemm_f32: NCHW needs M %% hw == 0, no A2");
    else
        PETR_REQUIRE(lda >= K, PETR_E_BADARG, "petr_gemm_f32: lda < K");
    if (A2) PETR_REQUIRE(a2_rows > 0, PETR_E_BADARG, "petr_gemm_f32: a2_rows <= 0");
    if (act == PETR_ACT_SIGMOID_GATE) PETR_REQUIRE(gate, PETR_E_BADARG, "petr_gemm_f32: gate is null");
    // the 256 -> <= 10 tails of the branches: streaming kernel with W in registers
    if (tuning::gemm_skinny() && a_layout == PETR_A_ROWMAJOR && K == 256 && N <= SKINNY_MAX_N && M >= 4096 && !A2 && !gate &&
        !C_bf16 && C && (act == PETR_ACT_NONE || act == PETR_ACT_RELU) && (lda % 4) == 0 && aligned16(A) &&
        aligned16(W)) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int ctas = (M + 7) / 8 < sms * 8 ? (M + 7) / 8 : sms * 8;
        gemm_f32_skinny_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(A, lda, W, bias, residual, C, ldc, M, N,
                                                                     act == PETR_ACT_RELU ? 1 : 0);
        return after_launch("petr_gemm_f32");
    }
    GemmParams p{A, lda, a_layout, hw, A2, a2_rows, W, bias, residual, gate, act, C,
                 (__nv_bfloat16*)C_bf16, ldc, M, N, K};
    dim3 grid((M + GBM - 1) / GBM, (N + GBN - 1) / GBN);
    gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return after_launch("petr_gemm_f32");
}

extern "C" int petr_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V,
                                  int ldv, const uint8_t* key_mask, const float* attn_bias, int bias_per_head,
                                  float* O, int ldo, void* O_split, float* lse, float drop_p, uint64_t seed,
                                  float* partial, int ksplit,
                                  int B, int Lq, int Lk, int heads, int head_dim, float scale, void* stream) {
    PETR_REQUIRE(Q && K && V && (O || O_split), PETR_E_BADARG, "petr_attention_f32: null pointer");
    PETR_REQUIRE(B > 0 && Lq > 0 && Lk > 0 && heads > 0, PETR_E_BADARG, "petr_attention_f32: bad dimension");
    PETR_REQUIRE(head_dim == 32, PETR_E_SHAPE, "petr_attention_f32: head_dim must be 32, got %d", head_dim);
    PETR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, PETR_E_BADARG, "petr_attention_f32: drop_p must be in [0, 1)");
    PETR_REQUIRE((ldk & 3) == 0 && (ldv & 3) == 0 && aligned16(K) && aligned16(V) && (ldq & 3) == 0 && aligned16(Q),
                 PETR_E_ALIGN, "petr_attention_f32: Q/K/V rows must be 16-byte aligned");
    if (O) PETR_REQUIRE((ldo & 3) == 0 && aligned16(O), PETR_E_ALIGN, "petr_attention_f32: O rows must be 16-byte aligned");
    if (ksplit < 1) ksplit = 1;
    const int tiles = (Lk + ATT_KT - 1) / ATT_KT;
    if (ksplit > tiles) ksplit = tiles;
    PETR_REQUIRE(ksplit == 1 || (partial && aligned16(partial)), PETR_E_BADARG,
                 "petr_attention_f32: ksplit > 1 needs a 16-byte aligned partial buffer (ksplit * petr_attention_partial_floats)");
    PETR_REQUIRE((int64_t)B * ksplit <= 65535, PETR_E_SHAPE, "petr_attention_f32: B * ksplit too large");
    dim3 grid((Lq + ATT_ROWS - 1) / ATT_ROWS, heads, B * ksplit);
    attention_f32_kernel<<<grid, ATT_ROWS, 0, (cudaStream_t)stream>>>(
        Q, ldq, K, ldk, V, ldv, key_mask, O, ldo, (__nv_bfloat16*)O_split, heads * 32, Lq, Lk,
        scale * 1.4426950408889634f, attn_bias, bias_per_head ? (int64_t)Lq * Lk : 0, lse, drop_p, seed, partial, ksplit);
    int rc = after_launch("petr_attention_f32");
    if (rc || ksplit == 1) return rc;
    return petr_attention_merge(partial, ksplit, O, ldo, O_split, lse, B, Lq, heads, stream);
}

extern "C" int petr_layernorm_f32(const float* X, const float* gamma, const float* beta, float* Y,
                                  int rows, int E, float eps, int relu, int nan_to_num, void* stream) {
    PETR_REQUIRE(X && gamma && beta && Y, PETR_E_BADARG, "petr_layernorm_f32: null pointer");
    PETR_REQUIRE(rows > 0 && E > 0, PETR_E_BADARG, "petr_layernorm_f32: bad dimension");
    dim3 grid((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
#define PETR_LN_CASE(PER) case PER * 32: \
        layernorm_f32_kernel<PER><<<grid, 256, 0, st>>>(X, gamma, beta, Y, rows, eps, relu, nan_to_num); break;
    switch (E) {
        PETR_LN_CASE(1) PETR_LN_CASE(2) PETR_LN_CASE(4) PETR_LN_CASE(8)
        PETR_LN_CASE(16) PETR_LN_CASE(32) PETR_LN_CASE(64)
        default:
            return fail(PETR_E_SHAPE, "petr_layernorm_f32: E must be 32 * 2^k <= 2048, got %d", E);
    }
#undef PETR_LN_CASE
    return after_launch("petr_layernorm_f32");
}

extern "C" int petr_decode_boxes(float* reg, const float* ref, const float* h_pc_range,
                                 const float* time_div, int L, int B, int Q, int code_size, void* stream) {
    PETR_REQUIRE(reg && ref && h_pc_range, PETR_E_BADARG, "petr_decode_boxes: null pointer");
    PETR_REQUIRE(L > 0 && B > 0 && Q > 0 && code_size >= 5, PETR_E_BADARG, "petr_decode_boxes: bad dimension");
    DecodeParams dp;
    for (int i = 0; i < 6; ++i) dp.pc[i] = h_pc_range[i];
    int64_t rows = (int64_t)L * B * Q;
    decode_boxes_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reg, ref, dp, time_div, rows, B, Q, code_size);
    return after_launch("petr_decode_boxes");
}

extern "C" int petr_layernorm_multi_f32(const float* X, const float* gamma, const float* beta, float eps,
                                        const float* qpos, int qpos_rows, float* Y, void* Y_split,
                                        void* Ypos_split, const float* gamma2, const float* beta2,
                                        float* Z, void* Z_split, int rows, int E, int relu, int nparts,
                                        int64_t part_stride, void* stream) {
    PETR_REQUIRE(X && gamma && beta, PETR_E_BADARG, "petr_layernorm_multi_f32: null pointer");
    PETR_REQUIRE(nparts <= 1 || (part_stride > 0 && (part_stride % 4) == 0), PETR_E_BADARG,
                 "petr_layernorm_multi_f32: nparts > 1 needs a part_stride that is a multiple of 4 floats");
    PETR_REQUIRE(rows > 0, PETR_E_BADARG, "petr_layernorm_multi_f32: bad dimension");
    PETR_REQUIRE(E == 256, PETR_E_SHAPE, "petr_layernorm_multi_f32: E must be 256, got %d", E);
    if (Ypos_split) PETR_REQUIRE(qpos && qpos_rows > 0, PETR_E_BADARG, "petr_layernorm_multi_f32: qpos required");
    if (Z || Z_split) PETR_REQUIRE(gamma2 && beta2, PETR_E_BADARG, "petr_layernorm_multi_f32: gamma2/beta2 required");
    PETR_REQUIRE(aligned16(X) && aligned16(gamma) && aligned16(beta) && aligned16(qpos) && aligned16(Y) &&
                 aligned16(Y_split) && aligned16(Ypos_split) && aligned16(gamma2) && aligned16(beta2) &&
                 aligned16(Z) && aligned16(Z_split), PETR_E_ALIGN,
                 "petr_layernorm_multi_f32: every pointer must be 16-byte aligned");
    LnMultiParams p{X, gamma, beta, eps, qpos, qpos_rows, Y, (__nv_bfloat16*)Y_split,
                    (__nv_bfloat16*)Ypos_split, (Z || Z_split) ? gamma2 : nullptr, beta2, Z,
                    (__nv_bfloat16*)Z_split, rows, relu, nparts < 1 ? 1 : nparts, part_stride};
    cudaError_t e = launch_pdl(layernorm_multi_kernel, dim3((rows + 7) / 8), dim3(256), 0, (cudaStream_t)stream, p);
    if (e != cudaSuccess) return fail((int)e, "petr_layernorm_multi_f32: launch: %s", cudaGetErrorString(e));
    return after_launch("petr_layernorm_multi_f32");
}

extern "C" int petr_split_rows_bf16(const float* X, int ldx, const float* add, int add_rows, void* out,
                                    int rows, int E, int relu, void* stream) {
    PETR_REQUIRE(X && out, PETR_E_BADARG, "petr_split_rows_bf16: null pointer");
    PETR_REQUIRE(rows > 0 && E > 0 && ldx >= E, PETR_E_BADARG, "petr_split_rows_bf16: bad dimension");
    if (add) PETR_REQUIRE(add_rows > 0, PETR_E_BADARG, "petr_split_rows_bf16: add_rows <= 0");
    int64_t total = (int64_t)rows * E;
    if (E % 8 == 0 && ldx % 4 == 0 && aligned16(X) && aligned16(out) && (!add || aligned16(add)))
        split_rows8_kernel<<<(unsigned)((total / 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            X, ldx, add, add_rows, (__nv_bfloat16*)out, rows, E, relu);
    else
        split_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            X, ldx, add, add_rows, (__nv_bfloat16*)out, rows, E, relu);
    return after_launch("petr_split_rows_bf16");
}

extern "C" int petr_x3_weight_bf16(const float* W, int N, int K, int transpose, void* out, void* stream) {
    PETR_REQUIRE(W && out, PETR_E_BADARG, "petr_x3_weight_bf16: null pointer");
    PETR_REQUIRE(N > 0 && K > 0, PETR_E_BADARG, "petr_x3_weight_bf16: bad dimension");
    int64_t total = (int64_t)N * K;
    x3_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, N, K, transpose ? 1 : 0,
                                                                                        (__nv_bfloat16*)out);
    return after_launch("petr_x3_weight_bf16");
}

extern "C" int petr_add_rows_bf16(const float* X, const float* add, int add_rows, void* out, int rows, int E,
                                  void* stream) {
    PETR_REQUIRE(X && add && out, PETR_E_BADARG, "petr_add_rows_bf16: null pointer");
    PETR_REQUIRE(rows > 0 && E > 0 && add_rows > 0 && (E % 4) == 0, PETR_E_BADARG, "petr_add_rows_bf16: bad dimension");
    PETR_REQUIRE(aligned16(X) && aligned16(add) && aligned16(out), PETR_E_ALIGN, "petr_add_rows_bf16: alignment");
    int64_t total = (int64_t)rows * (E / 4);
    add_rows_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        X, add, add_rows, (__nv_bfloat16*)out, rows, E / 4);
    return after_launch("petr_add_rows_bf16");
}

extern "C" int petr_stage_small(const void* h_pinned, void* dst, int nbytes, void* stream) {
    PETR_REQUIRE(h_pinned && dst, PETR_E_BADARG, "petr_stage_small: null pointer");
    PETR_REQUIRE(nbytes > 0 && (nbytes % 4) == 0, PETR_E_BADARG, "petr_stage_small: nbytes must be a positive multiple of 4");
    int nwords = nbytes / 4;
    int blocks = (nwords + 255) / 256;
    if (blocks > 64) blocks = 64;
    stage_small_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)h_pinned, (uint32_t*)dst, nwords);
    return after_launch("petr_stage_small");
}
