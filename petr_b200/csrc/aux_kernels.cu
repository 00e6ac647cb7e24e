// Kernels either side of the main path (SURVEY.md section 8f row 4 and the reference branches no shipped config takes):
//   * 2-D positional encodings of the with_multiview=False branch (petr_head.py:403-409,:414-420):
//     mmdet SinePositionalEncoding (un-vendored, mmdet 2.24.1) and LearnedPositionalEncoding3D
//     (positional_encoding.py:137-160);
//   * CPFPN tail (cp_fpn.py:159-210): nearest top-down add and the 3x3 level-0 conv as im2col over token-major maps
//     feeding the tcgen05 / fp32 GEMMs, token-major <-> NCHW layout passes;
//   * NormedLinear row / column norms (mmdet NormedLinear, petr_head.py:231-232).
#include "common.cuh"

namespace petr {
namespace {

// ---- mmdet SinePositionalEncoding: mask uint8 [G,H,W] -> [G*H*W, 2F] = [pos_y | pos_x], channel j of a half:
// even j -> sin(v / dim_t[j]), odd j -> cos(v / dim_t[j]) (interleaved, unlike the 3-D variant).  One block per (g, h).
template <typename OutT>
__global__ void __launch_bounds__(256)
sine2d_features_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ dim_t, int H, int W, int F,
                       int normalize, float scale, float eps, float offset, OutT* __restrict__ out) {
    extern __shared__ float sm[];            // [2][W] values, then [F] dim_t
    float* vy = sm;
    float* vx = sm + W;
    float* sdt = sm + 2 * W;
    const int h = blockIdx.x % H, g = blockIdx.x / H;
    const uint8_t* mg = mask + (int64_t)g * H * W;
    for (int i = threadIdx.x; i < F; i += blockDim.x) sdt[i] = dim_t[i];
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        int ey = 0, ly = 0;
        for (int k = 0; k < H; ++k) {
            int nm = mg[(int64_t)k * W + w] ? 0 : 1;
            ly += nm;
            if (k <= h) ey += nm;
        }
        float fy = (float)ey;
        if (normalize) fy = __fmul_rn(__fdiv_rn(__fadd_rn(fy, offset), __fadd_rn((float)ly, eps)), scale);
        vy[w] = fy;
    }
    if (threadIdx.x == 0) {
        int acc = 0;
        const uint8_t* row = mg + (int64_t)h * W;
        for (int w = 0; w < W; ++w) { acc += row[w] ? 0 : 1; vx[w] = (float)acc; }
        if (normalize) {
            float last = __fadd_rn((float)acc, eps);
            for (int w = 0; w < W; ++w) vx[w] = __fmul_rn(__fdiv_rn(__fadd_rn(vx[w], offset), last), scale);
        }
    }
    __syncthreads();
    const int C = 2 * F;
    OutT* orow = out + ((int64_t)blockIdx.x * W) * C;
    for (int i = threadIdx.x; i < W * C; i += blockDim.x) {
        const int w = i / C, ch = i % C;
        const int j = ch % F;
        const float val = __fdiv_rn(ch < F ? vy[w] : vx[w], sdt[j]);
        const float r = (j & 1) ? cosf(val) : sinf(val);
        if constexpr (sizeof(OutT) == 4) orow[i] = r;
        else orow[i] = __float2bfloat16_rn(r);
    }
}

// ---- LearnedPositionalEncoding3D.forward: [G*H*W, 2F] = [col_embed[w] | row_embed[h]]
__global__ void learned_pos_features_kernel(const float* __restrict__ row_embed, const float* __restrict__ col_embed,
                                            int64_t total, int H, int W, int F, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % (2 * F));
    const int64_t t = i / (2 * F);
    const int w = (int)(t % W), h = (int)((t / W) % H);
    out[i] = ch < F ? col_embed[(int64_t)w * F + ch] : row_embed[(int64_t)h * F + (ch - F)];
}

// ---- token-major fp32 [G*hw, C] -> NCHW fp32 [G, C, hw] (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256)
tokens_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int hw) {
    __shared__ float tile[32][33];
    const int g = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int pp = p0 + i, c = c0 + tx;
        tile[i][tx] = (pp < hw && c < C) ? x[((int64_t)g * hw + pp) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, pp = p0 + tx;
        if (c < C && pp < hw) out[((int64_t)g * C + c) * hw + pp] = tile[tx][i];
    }
}

__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size, float scale) {
    // torch upsample_nearest (legacy 'nearest'): floor(dst * (in / out)) in fp32, clamped
    if (in_size == out_size) return dst;
    int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < in_size - 1 ? s : in_size - 1;
}

// ---- FPN top-down step on token-major maps: dst[g,h,w,:] += src[g, nearest(h), nearest(w), :]
__global__ void upsample_add_tokens_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t total4,
                                           int H, int W, int h, int w, int C4, float sh, float sw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)(i % C4);
    const int64_t t = i / C4;
    const int x = (int)(t % W), y = (int)((t / W) % H);
    const int64_t g = t / ((int64_t)W * H);
    const int ys = nearest_src(y, h, H, sh), xs = nearest_src(x, w, W, sw);
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[((g * h + ys) * w + xs) * C4 + c];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
}

// ---- im2col of a 3x3 / pad 1 / stride 1 conv over a token-major map: out[t, tap*C + c] = x[g, y+dy, x+dx, c]
// (tap = 3*(dy+1) + (dx+1), zero outside).  The conv weight is repacked once to [Cout][tap][Cin] by the caller.
// OutMode 0: fp32; 1: bf16; 2: bf16 [hi | lo] halves of width 9C each (bf16x3 GEMM operand).
template <int OutMode>
__global__ void im2col3x3_tokens_kernel(const float* __restrict__ x, void* __restrict__ out, int64_t total4,
                                        int H, int W, int C4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c = (int)(i % C4);
    const int tap = (int)((i / C4) % 9);
    const int64_t t = i / (9 * (int64_t)C4);
    const int xx = (int)(t % W), yy = (int)((t / W) % H);
    const int64_t g = t / ((int64_t)W * H);
    const int ys = yy + tap / 3 - 1, xs = xx + tap % 3 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ys >= 0 && ys < H && xs >= 0 && xs < W)
        v = __ldg(reinterpret_cast<const float4*>(x) + ((g * H + ys) * W + xs) * C4 + c);
    const int64_t K4 = 9 * (int64_t)C4;
    const int64_t col4 = (int64_t)tap * C4 + c;
    if constexpr (OutMode == 0) {
        reinterpret_cast<float4*>(out)[t * K4 + col4] = v;
    } else {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        uint2 hi = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        if constexpr (OutMode == 1) {
            reinterpret_cast<uint2*>(out)[t * K4 + col4] = hi;
        } else {
            __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
            __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
            reinterpret_cast<uint2*>(out)[t * 2 * K4 + col4] = hi;
            reinterpret_cast<uint2*>(out)[t * 2 * K4 + K4 + col4] = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
        }
    }
}

// ---- mmdet NormedLinear: x / (||x||_dim1 ^ power + eps) * temperature.  For the 3-D (B, Q, E) activations the head
// passes, dim 1 is the QUERY axis (the published module normalises dim=1 whatever the rank): one norm per (b, channel).
// Also used with B = 1, Q = rows for the weight (norm over dim 1 of a 2-D [N, K] weight = per row, handled by
// row_norm_scale_kernel).
__global__ void __launch_bounds__(256)
col_norm_scale_kernel(const float* __restrict__ x, float* __restrict__ out, int Q, int E, float power, float eps,
                      float mult) {
    // block = one (b, 32-channel group); 8 warps stride over the queries
    __shared__ float part[8][32];
    const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), wrp = threadIdx.x >> 5;
    const float* xb = x + (int64_t)b * Q * E;
    float acc = 0.f;
    if (c < E) for (int q = wrp; q < Q; q += 8) { float v = xb[(int64_t)q * E + c]; acc = fmaf(v, v, acc); }
    part[wrp][threadIdx.x & 31] = acc;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i][threadIdx.x & 31];
    const float inv = mult / (powf(sqrtf(tot), power) + eps);
    float* ob = out + (int64_t)b * Q * E;
    if (c < E) for (int q = wrp; q < Q; q += 8) ob[(int64_t)q * E + c] = xb[(int64_t)q * E + c] * inv;
}

__global__ void __launch_bounds__(256)
row_norm_scale_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int E, float power, float eps,
                      float mult) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (int64_t)row * E;
    float acc = 0.f;
    for (int i = lane; i < E; i += 32) acc = fmaf(xr[i], xr[i], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float inv = mult / (powf(sqrtf(acc), power) + eps);
    for (int i = lane; i < E; i += 32) out[(int64_t)row * E + i] = xr[i] * inv;
}

}  // namespace
}  // namespace petr

using namespace petr;

extern "C" int petr_sine2d_features(const uint8_t* mask, const float* dim_t, int G, int H, int W, int num_feats,
                                    int normalize, float scale, float eps, float offset, void* out, int out_dtype,
                                    void* stream) {
    PETR_REQUIRE(mask && dim_t && out, PETR_E_BADARG, "petr_sine2d_features: null pointer");
    PETR_REQUIRE(G > 0 && H > 0 && W > 0 && num_feats > 0, PETR_E_BADARG, "petr_sine2d_features: bad dimension");
    size_t smem = (size_t)(2 * W + num_feats) * sizeof(float);
    PETR_REQUIRE(smem <= 48 * 1024, PETR_E_SHAPE, "petr_sine2d_features: W too large");
    unsigned grid = (unsigned)(G * H);
    if (out_dtype == PETR_DTYPE_F32)
        sine2d_features_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(mask, dim_t, H, W, num_feats, normalize,
                                                                                 scale, eps, offset, (float*)out);
    else if (out_dtype == PETR_DTYPE_BF16)
        sine2d_features_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>(
            mask, dim_t, H, W, num_feats, normalize, scale, eps, offset, (__nv_bfloat16*)out);
    else
        return fail(PETR_E_BADARG, "petr_sine2d_features: bad dtype %d", out_dtype);
    return after_launch("petr_sine2d_features");
}

extern "C" int petr_learned_pos_features(const float* row_embed, int row_num, const float* col_embed, int col_num,
                                         int G, int H, int W, int num_feats, float* out, void* stream) {
    PETR_REQUIRE(row_embed && col_embed && out, PETR_E_BADARG, "petr_learned_pos_features: null pointer");
    PETR_REQUIRE(G > 0 && H > 0 && W > 0 && num_feats > 0, PETR_E_BADARG, "petr_learned_pos_features: bad dimension");
    PETR_REQUIRE(H <= row_num && W <= col_num, PETR_E_SHAPE,
                 "petr_learned_pos_features: map %dx%d exceeds the embedding tables (%d rows, %d cols)", H, W, row_num, col_num);
    const int64_t total = (int64_t)G * H * W * 2 * num_feats;
    learned_pos_features_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        row_embed, col_embed, total, H, W, num_feats, out);
    return after_launch("petr_learned_pos_features");
}

extern "C" int petr_tokens_to_nchw_f32(const float* x, float* out, int G, int C, int hw, void* stream) {
    PETR_REQUIRE(x && out, PETR_E_BADARG, "petr_tokens_to_nchw_f32: null pointer");
    PETR_REQUIRE(G > 0 && C > 0 && hw > 0 && G <= 65535, PETR_E_BADARG, "petr_tokens_to_nchw_f32: bad dimension");
    dim3 grid((hw + 31) / 32, (C + 31) / 32, G);
    tokens_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, C, hw);
    return after_launch("petr_tokens_to_nchw_f32");
}

extern "C" int petr_upsample_add_tokens(float* dst, const float* src, int G, int H, int W, int h, int w, int C,
                                        void* stream) {
    PETR_REQUIRE(dst && src, PETR_E_BADARG, "petr_upsample_add_tokens: null pointer");
    PETR_REQUIRE(G > 0 && H > 0 && W > 0 && h > 0 && w > 0 && C > 0, PETR_E_BADARG, "petr_upsample_add_tokens: bad dimension");
    PETR_REQUIRE((C % 4) == 0 && aligned16(dst) && aligned16(src), PETR_E_ALIGN,
                 "petr_upsample_add_tokens: C %% 4 == 0 and 16-byte aligned maps required");
    const int64_t total4 = (int64_t)G * H * W * (C / 4);
    upsample_add_tokens_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        dst, src, total4, H, W, h, w, C / 4, (float)h / (float)H, (float)w / (float)W);
    return after_launch("petr_upsample_add_tokens");
}

extern "C" int petr_im2col3x3_tokens(const float* x, void* out, int out_mode, int G, int H, int W, int C, void* stream) {
    PETR_REQUIRE(x && out, PETR_E_BADARG, "petr_im2col3x3_tokens: null pointer");
    PETR_REQUIRE(G > 0 && H > 0 && W > 0 && C > 0, PETR_E_BADARG, "petr_im2col3x3_tokens: bad dimension");
    PETR_REQUIRE((C % 4) == 0 && aligned16(x) && aligned16(out), PETR_E_ALIGN,
                 "petr_im2col3x3_tokens: C %% 4 == 0 and 16-byte aligned maps required");
    const int64_t total4 = (int64_t)G * H * W * 9 * (C / 4);
    const unsigned grid = (unsigned)((total4 + 255) / 256);
    if (out_mode == 0) im2col3x3_tokens_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, total4, H, W, C / 4);
    else if (out_mode == 1) im2col3x3_tokens_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, total4, H, W, C / 4);
    else if (out_mode == 2) im2col3x3_tokens_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, total4, H, W, C / 4);
    else return fail(PETR_E_BADARG, "petr_im2col3x3_tokens: bad out_mode %d", out_mode);
    return after_launch("petr_im2col3x3_tokens");
}

extern "C" int petr_norm_scale_f32(const float* x, float* out, int B, int Q, int E, int along_queries, float power,
                                   float eps, float mult, void* stream) {
    PETR_REQUIRE(x && out, PETR_E_BADARG, "petr_norm_scale_f32: null pointer");
    PETR_REQUIRE(B > 0 && Q > 0 && E > 0 && B <= 65535, PETR_E_BADARG, "petr_norm_scale_f32: bad dimension");
    if (along_queries) {
        dim3 grid((E + 31) / 32, B);
        col_norm_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, Q, E, power, eps, mult);
    } else {
        const int rows = B * Q;
        row_norm_scale_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, out, rows, E, power, eps, mult);
    }
    return after_launch("petr_norm_scale_f32");
}
