// Training path of the head: backward kernels for the fp32 forward kernels of
// simt_kernels.cu, the loss of petr_head.py:578-728 (sigmoid focal loss + weighted L1 on Hungarian-matched
// queries) and the matching-cost matrix of hungarian_assigner_3d.py:117-123.  First correct version: exact fp32
// FFMA arithmetic everywhere (the parity path); the tcgen05 backward is listed as future work in DESIGN.md.
//
//   forward op (kernel)                     backward here
//   petr_gemm_f32   C = (A+A2) W^T + b      dA = dC W            -> petr_gemm_f32 on a transposed W (caller)
//                                           dW = dC^T (A+A2)     -> petr_gemm_f32_tn (split over rows, atomics)
//                                           db = colsum(dC)      -> petr_colsum_f32
//                                           dA2 = periodic row sum of dA -> petr_rowsum_periodic_f32
//   ReLU / sigmoid gate epilogues           petr_relu_bwd_f32 / petr_sigmoid_gate_bwd_f32
//   petr_layernorm_f32                      petr_layernorm_bwd_f32 (dx, dgamma, dbeta)
//   petr_attention_f32 (+ lse, dropout)     petr_attention_f32_bwd (dQ kernel over queries, dK/dV kernel over keys)
//   petr_pos2posemb3d                       petr_pos2posemb3d_bwd
//   petr_decode_boxes                       petr_decode_boxes_bwd (d reg, d reference_points)
//   nn.Dropout                              petr_dropout_f32 (same kernel both ways: counter-based mask)
//   petr_norm_scale_f32 (NormedLinear)      petr_norm_scale_bwd_f32
//   petr_learned_pos_features               petr_learned_pos_bwd (d row_embed, d col_embed)
#include "common.cuh"

namespace petr {
namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------
// dW[n,k] += sum_m dY[m,n] * (A[m,k] + A2[m % a2_rows, k]).  Both operands are read along their contiguous axis
// (n resp. k); a CTA owns a 64 x 64 tile of dW and a chunk of rows, 16 x 16 threads x (4 x 4) outputs, partial sums
// merged with atomicAdd (dW is zero-initialised or holds a running gradient).
// ------------------------------------------------------------------------------------------
struct TnParams {
    const float* dY; int ldy; const float* A; int lda; int a_layout; int hw; const float* A2; int a2_rows;
    float* dW; int M, N, K, rows_per_cta;
};

__global__ void __launch_bounds__(256)
gemm_f32_tn_kernel(TnParams p) {
    __shared__ __align__(16) float sY[32][64 + 4];
    __shared__ __align__(16) float sA[32][64 + 4];
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int m_begin = blockIdx.z * p.rows_per_cta, m_end = min(p.M, m_begin + p.rows_per_cta);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int m0 = m_begin; m0 < m_end; m0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 64; i += 256) {
            const int r = i >> 6, c = i & 63, m = m0 + r;
            float y = 0.f, a = 0.f;
            if (m < m_end) {
                if (n0 + c < p.N) y = p.dY[(int64_t)m * p.ldy + n0 + c];
                if (k0 + c < p.K) {
                    const int k = k0 + c;
                    if (p.a_layout == PETR_A_NCHW) {
                        const int g = m / p.hw, rr = m - g * p.hw;
                        a = p.A[((int64_t)g * p.K + k) * p.hw + rr];
                    } else {
                        a = p.A[(int64_t)m * p.lda + k];
                    }
                    if (p.A2) a += p.A2[(int64_t)(m % p.a2_rows) * p.lda + k];
                }
            }
            sY[r][c] = y;
            sA[r][c] = a;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float4 y4 = *reinterpret_cast<const float4*>(&sY[r][ty * 4]);
            const float4 a4 = *reinterpret_cast<const float4*>(&sA[r][tx * 4]);
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yv[i], av[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= p.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < p.K) atomicAdd(p.dW + (int64_t)n * p.K + k, acc[i][j]);
        }
    }
}

// out[n] += sum_m X[m, n]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int ldx, int M, int N, int rows_per_cta, float* __restrict__ out) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    const int m_begin = blockIdx.y * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
    __shared__ float part[8][32];
    float acc = 0.f;
    if (n < N) for (int m = m_begin + w; m < m_end; m += 8) acc += X[(int64_t)m * ldx + n];
    part[w][threadIdx.x & 31] = acc;
    __syncthreads();
    if (w == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(out + n, t);
    }
}

// out[r, c] += sum_j X[j * rows + r, c]   (gradient of a row-periodic broadcast add)
__global__ void rowsum_periodic_kernel(const float* __restrict__ X, int64_t total, int period_elems, int reps,
                                       float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= period_elems) return;
    float acc = 0.f;
    for (int j = 0; j < reps; ++j) acc += X[(int64_t)j * period_elems + i];
    out[i] += acc;
}

__global__ void relu_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, float* __restrict__ dX,
                                int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dX[i] = Y[i] > 0.f ? dY[i] : 0.f;
}

// out = gate * sigmoid(z) (petr_gemm_f32 PETR_ACT_SIGMOID_GATE; z itself is not kept): given out, gate and d out,
// sigmoid = out / gate;  dz = dout * gate * s (1 - s),  dgate = dout * s.  s is recomputed from z_pre when supplied.
__global__ void sigmoid_gate_bwd_kernel(const float* __restrict__ dOut, const float* __restrict__ Zpre,
                                        const float* __restrict__ gate, float* __restrict__ dZ,
                                        float* __restrict__ dGate, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = 1.f / (1.f + expf(-Zpre[i]));
    const float g = dOut[i];
    dZ[i] = g * gate[i] * s * (1.f - s);
    dGate[i] = g * s;
}

__global__ void dropout_kernel(const float* __restrict__ X, const float* __restrict__ R, float* __restrict__ Y,
                               int64_t n, float p, float scale, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Y[i] = (dropout_keep(seed, (uint64_t)i, p) ? X[i] * scale : 0.f) + (R ? R[i] : 0.f);
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row (E = 32 * PER).  y = xhat * gamma + beta (then optional ReLU / nan_to_num:
// their masks are re-derived from the recomputed y).  dgamma / dbeta: per-CTA reduction, then atomics.
// ------------------------------------------------------------------------------------------
template <int PER>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ X, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ dY, float* __restrict__ dX, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int rows, float eps, int relu, int nan_to_num) {
    constexpr int E = PER * 32;
    __shared__ float sg[8][E], sb[8][E];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + w;
    float dg[PER], db[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) dg[i] = db[i] = 0.f;
    if (row < rows) {
        const float* x = X + (int64_t)row * E;
        float v[PER], s = 0.f;
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
        const bool finite_row = isfinite(mean) && isfinite(rstd);
        float g[PER], xh[PER], sg1 = 0.f, sg2 = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            xh[i] = (v[i] - mean) * rstd;
            const float y = xh[i] * gamma[c] + beta[c];
            float d = dY[(int64_t)row * E + c];
            if (relu && !(y > 0.f)) d = 0.f;
            if (nan_to_num && !finite_row) d = 0.f;        // NaN rows (fully masked sample) were replaced by constants
            dg[i] = finite_row ? d * xh[i] : 0.f;
            db[i] = d;
            g[i] = d * gamma[c];
            sg1 += g[i];
            sg2 += g[i] * xh[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sg1 += __shfl_xor_sync(0xffffffffu, sg1, o);
            sg2 += __shfl_xor_sync(0xffffffffu, sg2, o);
        }
        sg1 /= (float)E;
        sg2 /= (float)E;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const float r = rstd * (g[i] - sg1 - xh[i] * sg2);
            dX[(int64_t)row * E + lane + 32 * i] = finite_row ? r : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) { sg[w][lane + 32 * i] = dg[i]; sb[w][lane + 32 * i] = db[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < E; c += 256) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += sg[i][c]; b += sb[i][c]; }
        if (dgamma) atomicAdd(dgamma + c, a);
        if (dbeta) atomicAdd(dbeta + c, b);
    }
}

// ------------------------------------------------------------------------------------------
// Attention backward (head_dim 32), flash style: nothing of size Lq x Lk is stored; P is recomputed from the
// forward's log-sum-exp.  S = scale q.k (+ bias), P = softmax(S), Pd = dropout(P), O = Pd V.
//   D_q = dO_q . O_q;  dPd = dO V^T;  dS = P o (mask/(1-p) dPd - D);  dQ = scale dS K;  dK = scale dS^T Q;  dV = Pd^T dO
// Two kernels so that neither needs atomics over its own output rows: one thread per QUERY streaming key tiles from
// shared memory (dQ; optional key split with atomics), one thread per KEY streaming query tiles (dK, dV).
// ------------------------------------------------------------------------------------------
constexpr int AB_ROWS = 128, AB_TILE = 32;


struct AttBwdParams {
    const float* Q; int ldq; const float* K; int ldk; const float* V; int ldv; const float* O; int ldo;
    const float* dO; int lddo; const float* lse; const uint8_t* key_mask; const float* attn_bias; int64_t bias_bh_stride;
    float* dQ; int lddq; float* dK; int lddk; float* dV; int lddv;
    int Lq, Lk, heads; float scale; float drop_p; uint64_t seed; int ksplit;
};

__global__ void __launch_bounds__(AB_ROWS)
attention_bwd_dq_kernel(AttBwdParams p) {
    __shared__ __align__(16) float sKb[2][AB_TILE][32];
    __shared__ __align__(16) float sVb[2][AB_TILE][32];
    __shared__ float sBiasb[2][AB_TILE];
    const int head = blockIdx.y, b = blockIdx.z / p.ksplit, part = blockIdx.z % p.ksplit;
    const int q = blockIdx.x * AB_ROWS + threadIdx.x;
    const bool live = q < p.Lq;
    const int bh = b * p.heads + head;
    const float scale_log2 = p.scale * LOG2E;
    uint64_t q2[16], do2[16], dq2[16];            // packed fp32 pairs (fma.rn.f32x2: two FMAs per issued instruction)
    float D = 0.f, lse = 0.f;
    {
        const int64_t row = (int64_t)b * p.Lq + (live ? q : 0);
        const float* qp = p.Q + row * p.ldq + head * 32;
        const float* dop = p.dO + row * p.lddo + head * 32;
        const float* op = p.O + row * p.ldo + head * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            q2[i >> 1] = pack2(qp[i] * scale_log2, qp[i + 1] * scale_log2);
            do2[i >> 1] = pack2(dop[i], dop[i + 1]);
            D = fmaf(dop[i], op[i], D);
            D = fmaf(dop[i + 1], op[i + 1], D);
            dq2[i >> 1] = 0ull;
        }
        lse = p.lse[(int64_t)bh * p.Lq + (live ? q : 0)];
    }
    const bool dead_row = !live || !isfinite(lse) || !isfinite(D);      // fully masked row: NaN forward, zero gradient
    const float* brow = p.attn_bias ? p.attn_bias + (int64_t)bh * p.bias_bh_stride + (int64_t)(live ? q : 0) * p.Lk : nullptr;
    const uint64_t drop_row = ((uint64_t)bh * p.Lq + (live ? q : 0)) * (uint64_t)p.Lk;
    const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const float* Kb = p.K + (int64_t)b * p.Lk * p.ldk + head * 32;
    const float* Vb = p.V + (int64_t)b * p.Lk * p.ldv + head * 32;
    const int tiles = (p.Lk + AB_TILE - 1) / AB_TILE;
    const int t_begin = (int)((int64_t)tiles * part / p.ksplit), t_end = (int)((int64_t)tiles * (part + 1) / p.ksplit);
    auto issue_tile = [&](int t, int buf) {
        const int k0 = t * AB_TILE;
        for (int i = threadIdx.x; i < AB_TILE * 8; i += AB_ROWS) {
            const int r = i >> 3, c4 = (i & 7) * 4;
            const bool valid = k0 + r < p.Lk;
            const int64_t row = valid ? k0 + r : 0;
            cp_async16(&sKb[buf][r][c4], Kb + row * p.ldk + c4, valid);
            cp_async16(&sVb[buf][r][c4], Vb + row * p.ldv + c4, valid);
        }
        if (threadIdx.x < AB_TILE) {
            const int r = threadIdx.x;
            const bool dead = (k0 + r >= p.Lk) || (p.key_mask && p.key_mask[(int64_t)b * p.Lk + k0 + r]);
            sBiasb[buf][r] = dead ? -INFINITY : 0.f;
        }
        cp_async_commit();
    };
    if (t_begin < t_end) issue_tile(t_begin, 0);
    for (int t = t_begin; t < t_end; ++t) {
        const int k0 = t * AB_TILE, buf = (t - t_begin) & 1;
        if (t + 1 < t_end) { issue_tile(t + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();                                   // tile t has landed for every thread
        const float (*sK)[32] = sKb[buf];
        const float (*sV)[32] = sVb[buf];
        const float* sBias = sBiasb[buf];
        if (!dead_row) {
#pragma unroll 2
        for (int r = 0; r < AB_TILE; ++r) {
            uint64_t s0 = 0ull, s1 = 0ull, s2 = 0ull, s3 = 0ull, d0 = 0ull, d1 = 0ull, d2 = 0ull, d3 = 0ull;   // short chains
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const ulonglong2 ka = *reinterpret_cast<const ulonglong2*>(&sK[r][2 * i]);
                const ulonglong2 kb = *reinterpret_cast<const ulonglong2*>(&sK[r][2 * i + 4]);
                const ulonglong2 va = *reinterpret_cast<const ulonglong2*>(&sV[r][2 * i]);
                const ulonglong2 vb = *reinterpret_cast<const ulonglong2*>(&sV[r][2 * i + 4]);
                s0 = fma2(q2[i], ka.x, s0); s1 = fma2(q2[i + 1], ka.y, s1); s2 = fma2(q2[i + 2], kb.x, s2); s3 = fma2(q2[i + 3], kb.y, s3);
                d0 = fma2(do2[i], va.x, d0); d1 = fma2(do2[i + 1], va.y, d1); d2 = fma2(do2[i + 2], vb.x, d2); d3 = fma2(do2[i + 3], vb.y, d3);
            }
            const float2 su = unpack2(s0), sw = unpack2(s1), sy = unpack2(s2), sz = unpack2(s3);
            const float2 du = unpack2(d0), dw = unpack2(d1), dy = unpack2(d2), dz = unpack2(d3);
            float s = (((su.x + su.y) + (sw.x + sw.y)) + ((sy.x + sy.y) + (sz.x + sz.y))) + sBias[r];
            if (brow && k0 + r < p.Lk) s += brow[k0 + r] * LOG2E;
            const float pr = exp2f(s - lse);                                  // -inf -> 0
            float dp = ((du.x + du.y) + (dw.x + dw.y)) + ((dy.x + dy.y) + (dz.x + dz.y));
            if (p.drop_p > 0.f) dp = dropout_keep(p.seed, drop_row + (uint64_t)(k0 + r), p.drop_p) ? dp * drop_scale : 0.f;
            const float ds = pr * (dp - D);
            const uint64_t ds2 = pack2(ds, ds);
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(&sK[r][2 * i]);
                dq2[i] = fma2(ds2, kk.x, dq2[i]);
                dq2[i + 1] = fma2(ds2, kk.y, dq2[i + 1]);
            }
        }
        }
        __syncthreads();                                   // buffer free before tile t + 2 is issued into it
    }
    float dq[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float2 t = unpack2(dq2[i]); dq[2 * i] = t.x; dq[2 * i + 1] = t.y; }
    if (live) {
        float* out = p.dQ + ((int64_t)b * p.Lq + q) * p.lddq + head * 32;
        if (p.ksplit == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) out[i] = dq[i] * p.scale;
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(out + i, dq[i] * p.scale);
        }
    }
}

__global__ void __launch_bounds__(AB_ROWS)
attention_bwd_dkv_kernel(AttBwdParams p) {
    __shared__ __align__(16) float sQb[2][AB_TILE][32];
    __shared__ __align__(16) float sdOb[2][AB_TILE][32];
    __shared__ __align__(16) float sOb[2][AB_TILE][32];
    __shared__ float sLseRaw[2][AB_TILE];
    __shared__ float sLse[AB_TILE], sD[AB_TILE];
    const int head = blockIdx.y, b = blockIdx.z;
    const int k = blockIdx.x * AB_ROWS + threadIdx.x;
    const int bh = b * p.heads + head;
    const bool live = k < p.Lk && !(p.key_mask && p.key_mask[(int64_t)b * p.Lk + k]);
    const float scale_log2 = p.scale * LOG2E;
    uint64_t k2[16], v2[16], dk2[16], dv2[16];    // packed fp32 pairs; k2 carries scale * log2(e)
    {
        const int64_t row = (int64_t)b * p.Lk + (k < p.Lk ? k : 0);
        const float* kp = p.K + row * p.ldk + head * 32;
        const float* vp = p.V + row * p.ldv + head * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            k2[i >> 1] = pack2(kp[i] * scale_log2, kp[i + 1] * scale_log2);
            v2[i >> 1] = pack2(vp[i], vp[i + 1]);
            dk2[i >> 1] = 0ull;
            dv2[i >> 1] = 0ull;
        }
    }
    const float drop_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const float* Qb = p.Q + (int64_t)b * p.Lq * p.ldq + head * 32;
    const float* dOb = p.dO + (int64_t)b * p.Lq * p.lddo + head * 32;
    const float* Ob = p.O + (int64_t)b * p.Lq * p.ldo + head * 32;
    const float* lse_b = p.lse + (int64_t)bh * p.Lq;
    auto issue_tile = [&](int q0, int buf) {
        for (int i = threadIdx.x; i < AB_TILE * 8; i += AB_ROWS) {
            const int r = i >> 3, c4 = (i & 7) * 4;
            const bool valid = q0 + r < p.Lq;
            const int64_t row = valid ? q0 + r : 0;
            cp_async16(&sQb[buf][r][c4], Qb + row * p.ldq + c4, valid);
            cp_async16(&sdOb[buf][r][c4], dOb + row * p.lddo + c4, valid);
            cp_async16(&sOb[buf][r][c4], Ob + row * p.ldo + c4, valid);
        }
        if (threadIdx.x < AB_TILE) {
            const bool valid = q0 + (int)threadIdx.x < p.Lq;
            cp_async4(&sLseRaw[buf][threadIdx.x], lse_b + (valid ? q0 + (int)threadIdx.x : 0), valid);
        }
        cp_async_commit();
    };
    issue_tile(0, 0);
    for (int q0 = 0, it = 0; q0 < p.Lq; q0 += AB_TILE, ++it) {
        const int buf = it & 1;
        if (q0 + AB_TILE < p.Lq) { issue_tile(q0 + AB_TILE, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();                                   // tile has landed for every thread
        {
            // D = <dO, O> per query row: four lanes per row, eight columns each
            const int r = threadIdx.x >> 2, part = (threadIdx.x & 3) * 8;
            float D = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) D = fmaf(sdOb[buf][r][part + i], sOb[buf][r][part + i], D);
            D += __shfl_xor_sync(0xffffffffu, D, 1);
            D += __shfl_xor_sync(0xffffffffu, D, 2);
            if ((threadIdx.x & 3) == 0) {
                float l = sLseRaw[buf][r];
                if (q0 + r >= p.Lq || !isfinite(l) || !isfinite(D)) { l = INFINITY; D = 0.f; }   // exp2(s - inf) = 0: beyond Lq,
                sLse[r] = l;                                                                     // or a fully masked row
                sD[r] = D;
            }
        }
        __syncthreads();
        const float (*sQ)[32] = sQb[buf];
        const float (*sdO)[32] = sdOb[buf];
        if (live) {
#pragma unroll 2
        for (int r = 0; r < AB_TILE; ++r) {
            uint64_t s0 = 0ull, s1 = 0ull, s2 = 0ull, s3 = 0ull, d0 = 0ull, d1 = 0ull, d2 = 0ull, d3 = 0ull;   // short chains
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const ulonglong2 qa = *reinterpret_cast<const ulonglong2*>(&sQ[r][2 * i]);
                const ulonglong2 qb = *reinterpret_cast<const ulonglong2*>(&sQ[r][2 * i + 4]);
                const ulonglong2 da = *reinterpret_cast<const ulonglong2*>(&sdO[r][2 * i]);
                const ulonglong2 db = *reinterpret_cast<const ulonglong2*>(&sdO[r][2 * i + 4]);
                s0 = fma2(k2[i], qa.x, s0); s1 = fma2(k2[i + 1], qa.y, s1); s2 = fma2(k2[i + 2], qb.x, s2); s3 = fma2(k2[i + 3], qb.y, s3);
                d0 = fma2(v2[i], da.x, d0); d1 = fma2(v2[i + 1], da.y, d1); d2 = fma2(v2[i + 2], db.x, d2); d3 = fma2(v2[i + 3], db.y, d3);
            }
            const float2 su = unpack2(s0), sw = unpack2(s1), sy = unpack2(s2), sz = unpack2(s3);
            const float2 du = unpack2(d0), dw = unpack2(d1), dy = unpack2(d2), dz = unpack2(d3);
            float s = ((su.x + su.y) + (sw.x + sw.y)) + ((sy.x + sy.y) + (sz.x + sz.y));
            if (p.attn_bias && q0 + r < p.Lq)
                s += p.attn_bias[(int64_t)bh * p.bias_bh_stride + (int64_t)(q0 + r) * p.Lk + k] * LOG2E;
            const float pr = exp2f(s - sLse[r]);
            float m = 1.f;
            if (p.drop_p > 0.f)
                m = dropout_keep(p.seed, ((uint64_t)bh * p.Lq + (uint64_t)(q0 + r)) * (uint64_t)p.Lk + (uint64_t)k, p.drop_p)
                        ? drop_scale : 0.f;
            const float pd = pr * m;
            const float ds = pr * ((((du.x + du.y) + (dw.x + dw.y)) + ((dy.x + dy.y) + (dz.x + dz.y))) * m - sD[r]);
            const uint64_t pd2 = pack2(pd, pd), ds2 = pack2(ds, ds);
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const ulonglong2 dd = *reinterpret_cast<const ulonglong2*>(&sdO[r][2 * i]);
                const ulonglong2 qq = *reinterpret_cast<const ulonglong2*>(&sQ[r][2 * i]);
                dv2[i] = fma2(pd2, dd.x, dv2[i]); dv2[i + 1] = fma2(pd2, dd.y, dv2[i + 1]);
                dk2[i] = fma2(ds2, qq.x, dk2[i]); dk2[i + 1] = fma2(ds2, qq.y, dk2[i + 1]);
            }
        }
        }
        __syncthreads();                                   // buffers and sLse / sD free for the next tile
    }
    float dk[32], dv[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 a = unpack2(dk2[i]), c = unpack2(dv2[i]);
        dk[2 * i] = a.x; dk[2 * i + 1] = a.y; dv[2 * i] = c.x; dv[2 * i + 1] = c.y;
    }
    if (k < p.Lk) {
        float* ok = p.dK + ((int64_t)b * p.Lk + k) * p.lddk + head * 32;
        float* ov = p.dV + ((int64_t)b * p.Lk + k) * p.lddv + head * 32;
        // dK = scale * sum ds q (the tiles hold the raw q; scale * log2(e) rides on k2 for the scores)
#pragma unroll
        for (int i = 0; i < 32; ++i) { ok[i] = live ? dk[i] * p.scale : 0.f; ov[i] = live ? dv[i] : 0.f; }
    }
}

// ------------------------------------------------------------------------------------------
// small pieces
// ------------------------------------------------------------------------------------------
// pos2posemb3d: out[q, blk*F + j] = (j odd ? cos : sin)(2 pi ref[q, axis(blk)] / dim_t[j]), axis = (1, 0, 2)
__global__ void pos2posemb3d_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ dim_t,
                                        const float* __restrict__ dOut, int Q, int F, float* __restrict__ dref) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= Q * 3) return;
    const int q = w / 3, blk = w % 3;
    const int axis = blk == 0 ? 1 : (blk == 1 ? 0 : 2);
    const float pp = ref[q * 3 + axis] * 6.283185307179586f;
    float acc = 0.f;
    for (int j = lane; j < F; j += 32) {
        const float a = pp / dim_t[j];
        const float d = (j & 1) ? -sinf(a) : cosf(a);
        acc = fmaf(dOut[(int64_t)q * 3 * F + blk * F + j] * d, 6.283185307179586f / dim_t[j], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(dref + q * 3 + axis, acc);
}

// decode_boxes: out[0,1,4] = sigmoid(reg + inverse_sigmoid(ref)) * span + lo; out[8:] = reg / time_div (PETRv2)
struct DecodeBwdParams { float lo[3], span[3]; };
__global__ void decode_boxes_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dOut,
                                        const float* __restrict__ ref, DecodeBwdParams dp,
                                        const float* __restrict__ time_div, float* __restrict__ dReg,
                                        float* __restrict__ dref, int64_t rows, int B, int Q, int cs) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int q = (int)(r % Q), b = (int)((r / Q) % B);
    const int ch[3] = {0, 1, 4};
    for (int c = 0; c < cs; ++c) dReg[r * cs + c] = dOut[r * cs + c];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float s = (out[r * cs + ch[a]] - dp.lo[a]) / dp.span[a];
        const float dz = dOut[r * cs + ch[a]] * dp.span[a] * s * (1.f - s);
        dReg[r * cs + ch[a]] = dz;
        if (dref) {
            // inverse_sigmoid(x) = log(max(clamp(x,0,1), eps) / max(1 - clamp(x,0,1), eps)), eps = 1e-5 (autograd of it)
            const float x = ref[q * 3 + a];
            float g = 0.f;
            if (x >= 0.f && x <= 1.f) {
                if (x >= 1e-5f) g += 1.f / x;
                if (1.f - x >= 1e-5f) g += 1.f / (1.f - x);
            }
            atomicAdd(dref + q * 3 + a, dz * g);
        }
    }
    if (time_div) {
        const float d = time_div[b];
        for (int c = 8; c < cs; ++c) dReg[r * cs + c] = dOut[r * cs + c] / d;
    }
}

// ------------------------------------------------------------------------------------------
// loss (petr_head.py:578-645) and matching cost (hungarian_assigner_3d.py:117-123)
// ------------------------------------------------------------------------------------------
// normalize_bbox (core/bbox/util.py:38-61): (cx, cy, cz, w, l, h, rot[, vx, vy]) -> (cx, cy, log w, log l, cz, log h,
// sin, cos[, vx, vy])
__global__ void normalize_bbox_kernel(const float* __restrict__ g, int n, int cin, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* s = g + (int64_t)i * cin;
    float* o = out + (int64_t)i * (cin + 1);
    o[0] = s[0]; o[1] = s[1]; o[2] = logf(s[3]); o[3] = logf(s[4]); o[4] = s[2]; o[5] = logf(s[5]);
    o[6] = sinf(s[6]); o[7] = cosf(s[6]);
    for (int c = 7; c < cin; ++c) o[c + 1] = s[c];
}

// cost[q, g] = w_cls * FocalLossCost(cls[q], label_g) + w_reg * || pred[q, :8] - ngt[g, :8] ||_1, nan_to_num'ed as the
// assigner does (nan -> 100, +-inf -> +-100) before scipy's linear_sum_assignment on the host.
__global__ void hungarian_cost_kernel(const float* __restrict__ cls, const float* __restrict__ pred,
                                      const int64_t* __restrict__ labels, const float* __restrict__ ngt, int Q, int G,
                                      int C, int code, int gcode, float w_cls, float w_reg, float alpha, float gamma,
                                      float eps, float* __restrict__ cost) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * G) return;
    const int q = i / G, g = i % G;
    const float x = cls[(int64_t)q * C + (int)labels[g]];
    const float pr = 1.f / (1.f + expf(-x));
    const float neg = -logf(1.f - pr + eps) * (1.f - alpha) * powf(pr, gamma);
    const float pos = -logf(pr + eps) * alpha * powf(1.f - pr, gamma);
    float reg = 0.f;
    for (int d = 0; d < 8; ++d) reg += fabsf(pred[(int64_t)q * code + d] - ngt[(int64_t)g * gcode + d]);
    float c = (pos - neg) * w_cls + reg * w_reg;
    if (isnan(c)) c = 100.f;
    else if (isinf(c)) c = c > 0.f ? 100.f : -100.f;
    cost[i] = c;
}

// mmcv sigmoid_focal_loss (the op mmdet's FocalLoss calls on CUDA tensors) with per-row weights:
// loss += sum_{i,c} w_i FL(x_ic, [labels_i == c]); dX = scale * w_i dFL/dx (labels == C: background, no positive).
__global__ void __launch_bounds__(256)
focal_loss_kernel(const float* __restrict__ X, const int64_t* __restrict__ labels, const float* __restrict__ w,
                  int64_t n, int C, float alpha, float gamma, float grad_scale, float* __restrict__ loss_sum,
                  float* __restrict__ dX) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (i < n) {
        const int64_t row = i / C;
        const int c = (int)(i % C);
        const float x = X[i];
        const float pr = 1.f / (1.f + expf(-x));
        const float wt = w ? w[row] : 1.f;
        const bool pos = labels[row] == c;
        float d;
        if (pos) {
            const float t = powf(1.f - pr, gamma);
            const float lg = logf(fmaxf(pr, 1.17549435e-38f));
            l = -alpha * t * lg;
            d = alpha * t * (gamma * pr * lg - (1.f - pr));
        } else {
            const float t = powf(pr, gamma);
            const float lg = logf(fmaxf(1.f - pr, 1.17549435e-38f));
            l = -(1.f - alpha) * t * lg;
            d = (1.f - alpha) * t * (pr - gamma * (1.f - pr) * lg);
        }
        l *= wt;
        if (dX) dX[i] = d * wt * grad_scale;
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(loss_sum, t);
    }
}

// weighted L1 (mmdet L1Loss) over rows whose target is finite (petr_head.py:634-640):
// loss += sum w |x - t|;  dX = scale * w sign(x - t)
__global__ void __launch_bounds__(256)
l1_loss_kernel(const float* __restrict__ X, const float* __restrict__ T, const float* __restrict__ W, int64_t rows,
               int code, float grad_scale, float* __restrict__ loss_sum, float* __restrict__ dX) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (r < rows) {
        bool ok = true;
        for (int c = 0; c < code; ++c) ok = ok && isfinite(T[r * code + c]);
        for (int c = 0; c < code; ++c) {
            const float d = X[r * code + c] - T[r * code + c];
            const float wt = ok ? W[r * code + c] : 0.f;
            if (ok) l += wt * fabsf(d);
            if (dX) dX[r * code + c] = wt * grad_scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(loss_sum, t);
    }
}


// ---- petr_norm_scale_f32 backward.  y = x * s(n), s = mult / (n^power + eps), n = ||x|| over the group:
// dx = s * dy + x * (s'(n) / n) * sum(dy * x), s'(n) = -mult * power * n^(power-1) / (n^power + eps)^2;
// n == 0 contributes nothing through the norm (x is zero there; torch's norm backward masks it the same way).
__device__ __forceinline__ void norm_scale_coeffs(float sumsq, float dot, float power, float eps, float mult,
                                                  float& s, float& c) {
    const float n = sqrtf(sumsq);
    const float np = powf(n, power);
    const float den = np + eps;
    s = mult / den;
    c = n > 0.f ? -mult * power * powf(n, power - 2.f) / (den * den) * dot : 0.f;
}

__global__ void __launch_bounds__(256)
col_norm_scale_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int Q,
                          int E, float power, float eps, float mult) {
    // block = one (b, 32-channel group), as the forward: 8 warps stride over the queries
    __shared__ float p_sq[8][32], p_dot[8][32];
    const int b = blockIdx.y, lane = threadIdx.x & 31, c = blockIdx.x * 32 + lane, wrp = threadIdx.x >> 5;
    const int64_t base = (int64_t)b * Q * E;
    float sq = 0.f, dot = 0.f;
    if (c < E) for (int q = wrp; q < Q; q += 8) {
        const float v = x[base + (int64_t)q * E + c];
        sq = fmaf(v, v, sq);
        dot = fmaf(dy[base + (int64_t)q * E + c], v, dot);
    }
    p_sq[wrp][lane] = sq; p_dot[wrp][lane] = dot;
    __syncthreads();
    float tsq = 0.f, tdot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { tsq += p_sq[i][lane]; tdot += p_dot[i][lane]; }
    float s, cf;
    norm_scale_coeffs(tsq, tdot, power, eps, mult, s, cf);
    if (c < E) for (int q = wrp; q < Q; q += 8) {
        const int64_t i = base + (int64_t)q * E + c;
        dx[i] = fmaf(x[i], cf, s * dy[i]);
    }
}

__global__ void __launch_bounds__(256)
row_norm_scale_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int rows,
                          int E, float power, float eps, float mult) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (int64_t)row * E;
    const float* dr = dy + (int64_t)row * E;
    float sq = 0.f, dot = 0.f;
    for (int i = lane; i < E; i += 32) { sq = fmaf(xr[i], xr[i], sq); dot = fmaf(dr[i], xr[i], dot); }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    float s, cf;
    norm_scale_coeffs(sq, dot, power, eps, mult, s, cf);
    for (int i = lane; i < E; i += 32) dx[(int64_t)row * E + i] = fmaf(xr[i], cf, s * dr[i]);
}

// ---- petr_learned_pos_features backward: the token gradients [G*H*W, 2F] = [d col_embed[w] | d row_embed[h]] summed
// into the tables.  Block = one table row (W column rows first, then H row rows), thread = one feature channel.
__global__ void learned_pos_bwd_kernel(const float* __restrict__ d, int G, int H, int W, int F,
                                       float* __restrict__ d_row, float* __restrict__ d_col) {
    const int r = blockIdx.x;
    const int64_t ld = 2 * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float acc = 0.f;
        if (r < W) {
            for (int g = 0; g < G; ++g)
                for (int h = 0; h < H; ++h) acc += d[(((int64_t)g * H + h) * W + r) * ld + f];
            d_col[(int64_t)r * F + f] = acc;
        } else {
            const int h = r - W;
            for (int g = 0; g < G; ++g)
                for (int w = 0; w < W; ++w) acc += d[(((int64_t)g * H + h) * W + w) * ld + F + f];
            d_row[(int64_t)h * F + f] = acc;
        }
    }
}


// ---- operands of the weight-gradient GEMM on tcgen05: dW[n,k] = sum_m dY[m,n] (A[m,k] + A2[m % a2_rows, k]) is a
// K-major GEMM over m once both factors are transposed.  fp32 X[M, C] (+ add[m % add_rows, C]) -> bf16 out[C, w * Mpad]:
// w = 2: [hi | lo] (the A operand of a bf16x3 GEMM), w = 3: [hi | hi | lo] (its W operand); columns m >= M are zero.
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ add, int add_rows, int M, int C,
                       int Mpad, int w3, __nv_bfloat16* __restrict__ out) {
    // tile: 64 rows (m) x 32 columns (c); reads are 128-byte row segments, writes bf16 pairs = 128 bytes per warp
    __shared__ float tile[64][33];
    const int m0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 64; i += 8) {
        const int m = m0 + i, c = c0 + tx;
        float v = 0.f;
        if (m < M && c < C) {
            v = X[(int64_t)m * ldx + c];
            if (add) v += add[(int64_t)(m % add_rows) * C + c];
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    const int64_t ldo = (int64_t)(w3 ? 3 : 2) * Mpad;
    const int m = m0 + 2 * tx;                       // Mpad is even: m and m + 1 are inside or outside together
    if (m >= Mpad) return;
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i;
        if (c >= C) continue;
        const float v0 = tile[2 * tx][i], v1 = tile[2 * tx + 1][i];
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v0 - __low2float(hi), v1 - __high2float(hi));
        __nv_bfloat16* o = out + (int64_t)c * ldo + m;
        *reinterpret_cast<__nv_bfloat162*>(o) = hi;
        if (w3) {
            *reinterpret_cast<__nv_bfloat162*>(o + Mpad) = hi;
            *reinterpret_cast<__nv_bfloat162*>(o + 2 * (int64_t)Mpad) = lo;
        } else {
            *reinterpret_cast<__nv_bfloat162*>(o + Mpad) = lo;
        }
    }
}

// out[i] = sum_p parts[p][i] (the split-K parts of the weight-gradient GEMM), four elements per thread
__global__ void sum_parts_kernel(const float4* __restrict__ parts, int nparts, int64_t n4, float4* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = parts[i];
    for (int p = 1; p < nparts; ++p) {
        const float4 b = parts[(int64_t)p * n4 + i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
}

}  // namespace
}  // namespace petr

using namespace petr;

extern "C" int petr_gemm_f32_tn(const float* dY, int ldy, const float* A, int lda, int a_layout, int hw,
                                const float* A2, int a2_rows, float* dW, int M, int N, int K, void* stream) {
    PETR_REQUIRE(dY && A && dW, PETR_E_BADARG, "petr_gemm_f32_tn: null pointer");
    PETR_REQUIRE(M > 0 && N > 0 && K > 0, PETR_E_BADARG, "petr_gemm_f32_tn: bad dimension");
    PETR_REQUIRE(a_layout == PETR_A_ROWMAJOR || (a_layout == PETR_A_NCHW && hw > 0 && M % hw == 0), PETR_E_BADARG,
                 "petr_gemm_f32_tn: bad A layout");
    PETR_REQUIRE(!A2 || a2_rows > 0, PETR_E_BADARG, "petr_gemm_f32_tn: A2 needs a2_rows > 0");
    TnParams p{dY, ldy, A, lda, a_layout, hw, A2, a2_rows, dW, M, N, K, 0};
    const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
    int splits = (4 * 148 + tiles - 1) / tiles;                 // ~4 CTAs per SM in total
    const int max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.rows_per_cta = (((M + splits - 1) / splits) + 31) / 32 * 32;
    splits = (M + p.rows_per_cta - 1) / p.rows_per_cta;
    dim3 grid((N + 63) / 64, (K + 63) / 64, splits);
    gemm_f32_tn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return after_launch("petr_gemm_f32_tn");
}

extern "C" int petr_split_transpose_bf16(const float* X, int ldx, const float* add, int add_rows, int M, int C,
                                         int Mpad, int w3, void* out, void* stream) {
    PETR_REQUIRE(X && out, PETR_E_BADARG, "petr_split_transpose_bf16: null pointer");
    PETR_REQUIRE(M > 0 && C > 0 && ldx >= C && Mpad >= M && Mpad % 8 == 0, PETR_E_BADARG,
                 "petr_split_transpose_bf16: bad dimension");
    if (add) PETR_REQUIRE(add_rows > 0, PETR_E_BADARG, "petr_split_transpose_bf16: add_rows <= 0");
    dim3 grid((Mpad + 63) / 64, (C + 31) / 32);
    PETR_REQUIRE(grid.y <= 65535, PETR_E_SHAPE, "petr_split_transpose_bf16: C too large");
    split_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, add, add_rows, M, C, Mpad, w3 ? 1 : 0,
                                                                   (__nv_bfloat16*)out);
    return after_launch("petr_split_transpose_bf16");
}

extern "C" int petr_sum_parts_f32(const float* parts, int nparts, int64_t n, float* out, void* stream) {
    PETR_REQUIRE(parts && out, PETR_E_BADARG, "petr_sum_parts_f32: null pointer");
    PETR_REQUIRE(nparts > 0 && n > 0 && n % 4 == 0 && aligned16(parts) && aligned16(out), PETR_E_BADARG,
                 "petr_sum_parts_f32: n must be a positive multiple of 4, pointers 16-byte aligned");
    const int64_t n4 = n / 4;
    sum_parts_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)parts, nparts, n4,
                                                                                    (float4*)out);
    return after_launch("petr_sum_parts_f32");
}

extern "C" int petr_colsum_f32(const float* X, int ldx, int M, int N, float* out, void* stream) {
    PETR_REQUIRE(X && out, PETR_E_BADARG, "petr_colsum_f32: null pointer");
    PETR_REQUIRE(M > 0 && N > 0, PETR_E_BADARG, "petr_colsum_f32: bad dimension");
    int splits = (M + 255) / 256;                              // 900 query rows: 4 x N/32 CTAs; 24 000 tokens: 94 x N/32
    if (splits > 1024) splits = 1024;
    const int rows_per_cta = (M + splits - 1) / splits;
    dim3 grid((N + 31) / 32, (M + rows_per_cta - 1) / rows_per_cta);
    colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, M, N, rows_per_cta, out);
    return after_launch("petr_colsum_f32");
}

extern "C" int petr_rowsum_periodic_f32(const float* X, int rows, int E, int period_rows, float* out, void* stream) {
    PETR_REQUIRE(X && out, PETR_E_BADARG, "petr_rowsum_periodic_f32: null pointer");
    PETR_REQUIRE(rows > 0 && E > 0 && period_rows > 0 && rows % period_rows == 0, PETR_E_BADARG,
                 "petr_rowsum_periodic_f32: rows must be a multiple of period_rows");
    const int pe = period_rows * E;
    rowsum_periodic_kernel<<<(pe + 255) / 256, 256, 0, (cudaStream_t)stream>>>(X, (int64_t)rows * E, pe, rows / period_rows, out);
    return after_launch("petr_rowsum_periodic_f32");
}

extern "C" int petr_relu_bwd_f32(const float* dY, const float* Y, float* dX, int64_t n, void* stream) {
    PETR_REQUIRE(dY && Y && dX && n > 0, PETR_E_BADARG, "petr_relu_bwd_f32: bad argument");
    relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dY, Y, dX, n);
    return after_launch("petr_relu_bwd_f32");
}

extern "C" int petr_sigmoid_gate_bwd_f32(const float* dOut, const float* Zpre, const float* gate, float* dZ,
                                         float* dGate, int64_t n, void* stream) {
    PETR_REQUIRE(dOut && Zpre && gate && dZ && dGate && n > 0, PETR_E_BADARG, "petr_sigmoid_gate_bwd_f32: bad argument");
    sigmoid_gate_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dOut, Zpre, gate, dZ, dGate, n);
    return after_launch("petr_sigmoid_gate_bwd_f32");
}

extern "C" int petr_dropout_f32(const float* X, const float* residual, float* Y, int64_t n, float p, uint64_t seed,
                                void* stream) {
    PETR_REQUIRE(X && Y && n > 0 && p >= 0.f && p < 1.f, PETR_E_BADARG, "petr_dropout_f32: bad argument");
    dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, residual, Y, n, p, 1.f / (1.f - p), seed);
    return after_launch("petr_dropout_f32");
}

extern "C" int petr_layernorm_bwd_f32(const float* X, const float* gamma, const float* beta, const float* dY,
                                      float* dX, float* dgamma, float* dbeta, int rows, int E, float eps, int relu,
                                      int nan_to_num, void* stream) {
    PETR_REQUIRE(X && gamma && beta && dY && dX, PETR_E_BADARG, "petr_layernorm_bwd_f32: null pointer");
    PETR_REQUIRE(rows > 0, PETR_E_BADARG, "petr_layernorm_bwd_f32: bad dimension");
    PETR_REQUIRE(E == 256, PETR_E_SHAPE, "petr_layernorm_bwd_f32: E must be 256, got %d", E);
    layernorm_bwd_kernel<8><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(X, gamma, beta, dY, dX, dgamma, dbeta, rows,
                                                                             eps, relu, nan_to_num);
    return after_launch("petr_layernorm_bwd_f32");
}

extern "C" int petr_attention_f32_bwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                                      const float* O, int ldo, const float* dO, int lddo, const float* lse,
                                      const uint8_t* key_mask, const float* attn_bias, int bias_per_head,
                                      float* dQ, int lddq, float* dK, int lddk, float* dV, int lddv,
                                      float drop_p, uint64_t seed, int B, int Lq, int Lk, int heads, int head_dim,
                                      float scale, void* stream) {
    PETR_REQUIRE(Q && K && V && O && dO && lse && dQ && dK && dV, PETR_E_BADARG, "petr_attention_f32_bwd: null pointer");
    PETR_REQUIRE(B > 0 && Lq > 0 && Lk > 0 && heads > 0, PETR_E_BADARG, "petr_attention_f32_bwd: bad dimension");
    PETR_REQUIRE(head_dim == 32, PETR_E_SHAPE, "petr_attention_f32_bwd: head_dim must be 32");
    PETR_REQUIRE(((ldq | ldk | ldv | lddo | ldo) & 3) == 0 && aligned16(Q) && aligned16(K) && aligned16(V) &&
                     aligned16(dO) && aligned16(O),
                 PETR_E_ALIGN, "petr_attention_f32_bwd: rows must be 16-byte aligned");
    AttBwdParams p{Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, key_mask, attn_bias,
                   bias_per_head ? (int64_t)Lq * Lk : 0, dQ, lddq, dK, lddk, dV, lddv, Lq, Lk, heads, scale, drop_p, seed, 1};
    const int qblocks = (Lq + AB_ROWS - 1) / AB_ROWS;
    // 165 registers: three CTAs per SM.  Aim at two full waves of the CTAs that have work (a last query block with
    // fewer than half its rows live -- 900 = 7 x 128 + 4 -- ends at once and is not counted).
    const int busy = (Lq / AB_ROWS) + ((Lq % AB_ROWS) >= AB_ROWS / 2 ? 1 : 0);
    int ksplit = (2 * 3 * 148) / ((busy > 0 ? busy : 1) * heads * B);
    const int tiles = (Lk + AB_TILE - 1) / AB_TILE;
    if (ksplit > tiles / 8) ksplit = tiles / 8;
    if (ksplit < 1) ksplit = 1;
    p.ksplit = ksplit;
    if (ksplit > 1) {
        cudaError_t e = cudaMemset2DAsync(dQ, (size_t)lddq * 4, 0, (size_t)heads * 32 * 4, (size_t)B * Lq, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail((int)e, "petr_attention_f32_bwd: memset: %s", cudaGetErrorString(e));
    }
    dim3 gq(qblocks, heads, B * ksplit);
    attention_bwd_dq_kernel<<<gq, AB_ROWS, 0, (cudaStream_t)stream>>>(p);
    int rc = after_launch("petr_attention_f32_bwd(dq)");
    if (rc) return rc;
    dim3 gk((Lk + AB_ROWS - 1) / AB_ROWS, heads, B);
    attention_bwd_dkv_kernel<<<gk, AB_ROWS, 0, (cudaStream_t)stream>>>(p);
    return after_launch("petr_attention_f32_bwd(dkv)");
}

extern "C" int petr_norm_scale_bwd_f32(const float* x, const float* dOut, float* dX, int B, int Q, int E,
                                       int along_queries, float power, float eps, float mult, void* stream) {
    PETR_REQUIRE(x && dOut && dX, PETR_E_BADARG, "petr_norm_scale_bwd_f32: null pointer");
    PETR_REQUIRE(B > 0 && Q > 0 && E > 0 && B <= 65535, PETR_E_BADARG, "petr_norm_scale_bwd_f32: bad dimension");
    if (along_queries) {
        dim3 grid((E + 31) / 32, B);
        col_norm_scale_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, dOut, dX, Q, E, power, eps, mult);
    } else {
        const int rows = B * Q;
        row_norm_scale_bwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, dOut, dX, rows, E, power, eps, mult);
    }
    return after_launch("petr_norm_scale_bwd_f32");
}

extern "C" int petr_learned_pos_bwd(const float* dOut, int G, int H, int W, int num_feats, float* d_row_embed,
                                    float* d_col_embed, void* stream) {
    PETR_REQUIRE(dOut && d_row_embed && d_col_embed, PETR_E_BADARG, "petr_learned_pos_bwd: null pointer");
    PETR_REQUIRE(G > 0 && H > 0 && W > 0 && num_feats > 0, PETR_E_BADARG, "petr_learned_pos_bwd: bad dimension");
    const int threads = num_feats >= 256 ? 256 : ((num_feats + 31) / 32) * 32;
    learned_pos_bwd_kernel<<<(unsigned)(H + W), threads, 0, (cudaStream_t)stream>>>(dOut, G, H, W, num_feats,
                                                                                   d_row_embed, d_col_embed);
    return after_launch("petr_learned_pos_bwd");
}

extern "C" int petr_pos2posemb3d_bwd(const float* ref, const float* dim_t, const float* dOut, int Q, int num_feats,
                                     float* dref, void* stream) {
    PETR_REQUIRE(ref && dim_t && dOut && dref && Q > 0 && num_feats > 0, PETR_E_BADARG, "petr_pos2posemb3d_bwd: bad argument");
    const int threads = Q * 3 * 32;
    pos2posemb3d_bwd_kernel<<<(threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ref, dim_t, dOut, Q, num_feats, dref);
    return after_launch("petr_pos2posemb3d_bwd");
}

extern "C" int petr_decode_boxes_bwd(const float* out, const float* dOut, const float* ref, const float* h_pc_range,
                                     const float* time_div, float* dReg, float* dref, int L, int B, int Q, int code_size,
                                     void* stream) {
    PETR_REQUIRE(out && dOut && ref && h_pc_range && dReg, PETR_E_BADARG, "petr_decode_boxes_bwd: null pointer");
    PETR_REQUIRE(L > 0 && B > 0 && Q > 0 && code_size >= 5, PETR_E_BADARG, "petr_decode_boxes_bwd: bad dimension");
    DecodeBwdParams dp;
    for (int a = 0; a < 3; ++a) { dp.lo[a] = h_pc_range[a]; dp.span[a] = h_pc_range[3 + a] - h_pc_range[a]; }
    const int64_t rows = (int64_t)L * B * Q;
    decode_boxes_bwd_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        out, dOut, ref, dp, time_div, dReg, dref, rows, B, Q, code_size);
    return after_launch("petr_decode_boxes_bwd");
}

extern "C" int petr_normalize_bbox(const float* gt, int n, int code_in, float* out, void* stream) {
    PETR_REQUIRE(gt && out && n > 0 && code_in >= 7, PETR_E_BADARG, "petr_normalize_bbox: bad argument");
    normalize_bbox_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gt, n, code_in, out);
    return after_launch("petr_normalize_bbox");
}

extern "C" int petr_hungarian_cost(const float* cls, const float* pred, const int64_t* gt_labels, const float* ngt,
                                   int Q, int G, int C, int code, int gcode, float w_cls, float w_reg, float alpha,
                                   float gamma, float eps, float* cost, void* stream) {
    PETR_REQUIRE(cls && pred && gt_labels && ngt && cost, PETR_E_BADARG, "petr_hungarian_cost: null pointer");
    PETR_REQUIRE(Q > 0 && G > 0 && C > 0 && code >= 8 && gcode >= 8, PETR_E_BADARG, "petr_hungarian_cost: bad dimension");
    const int n = Q * G;
    hungarian_cost_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cls, pred, gt_labels, ngt, Q, G, C, code, gcode,
                                                                            w_cls, w_reg, alpha, gamma, eps, cost);
    return after_launch("petr_hungarian_cost");
}

extern "C" int petr_focal_loss(const float* X, const int64_t* labels, const float* row_weight, int64_t rows, int C,
                               float alpha, float gamma, float grad_scale, float* loss_sum, float* dX, void* stream) {
    PETR_REQUIRE(X && labels && loss_sum && rows > 0 && C > 0, PETR_E_BADARG, "petr_focal_loss: bad argument");
    const int64_t n = rows * C;
    focal_loss_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, labels, row_weight, n, C, alpha, gamma,
                                                                                     grad_scale, loss_sum, dX);
    return after_launch("petr_focal_loss");
}

extern "C" int petr_l1_loss(const float* X, const float* T, const float* W, int64_t rows, int code, float grad_scale,
                            float* loss_sum, float* dX, void* stream) {
    PETR_REQUIRE(X && T && W && loss_sum && rows > 0 && code > 0, PETR_E_BADARG, "petr_l1_loss: bad argument");
    l1_loss_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, T, W, rows, code, grad_scale, loss_sum, dX);
    return after_launch("petr_l1_loss");
}
