// NMS-free box decoding: NMSFreeCoder.decode_single
// (core/bbox/coders/nms_free_coder.py:48-101: sigmoid -> top-max_num over queries x classes -> labels / query
// index -> denormalize_bbox (core/bbox/util.py:63-87: atan2, exp) -> post_center_range and score filter) and the
// gravity-centre -> bottom-centre shift of PETRHead.get_bboxes (petr_head.py:731-751), one CTA per sample, one
// launch per batch: "features -> boxes" stays on the stream (CUDA-graph capturable, no host round trip).
//
// The top-k is a bitonic sort of (logit, flat index) keys in shared memory.  Ranking uses the LOGITS (sigmoid is
// monotonic), with the flat index ascending as tie-break, so labels and query indices are integer-exact and do not
// depend on the exp implementation; the returned scores are sigmoid(logit) in fp32.
#include "common.cuh"
#include <math.h>

namespace petr {
namespace {

constexpr int DEC_THREADS = 1024;

// float -> uint32 whose unsigned order equals the float order (NaN sorts above +inf; -0 < +0)
__device__ __forceinline__ uint32_t orderable(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct DecodeParams {
    const float* cls; const float* box;          // [B, Q, C], [B, Q, code]
    int Q, C, code, max_num, nsort, bottom_center, has_thr;
    float thr; float range[6];
    float* out_boxes; float* out_scores; int64_t* out_labels; int64_t* out_index; uint8_t* out_mask; int32_t* out_count;
};

__global__ void __launch_bounds__(DEC_THREADS, 1) nms_free_decode_kernel(DecodeParams p) {
    extern __shared__ unsigned long long keys[];            // nsort entries: (orderable(logit) << 32) | ~flat_index
    __shared__ int s_count;
    const int b = blockIdx.x;
    const int n = p.Q * p.C;
    const float* cls = p.cls + (int64_t)b * n;
    for (int i = threadIdx.x; i < p.nsort; i += DEC_THREADS) {
        unsigned long long k = 0ull;                         // padding sorts last
        if (i < n) k = ((unsigned long long)orderable(cls[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
        keys[i] = k;
    }
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // bitonic sort, descending
    for (int k = 2; k <= p.nsort; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < p.nsort; i += DEC_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], c = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int nb = p.code > 8 ? p.code - 1 : 7;              // decoded box width
    for (int r = threadIdx.x; r < p.max_num; r += DEC_THREADS) {
        const int64_t o = (int64_t)b * p.max_num + r;
        const unsigned long long k = keys[r];
        bool keep = r < n;
        float bx[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float score = 0.f;
        int64_t label = 0, qidx = 0;
        if (keep) {
            const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
            const float logit = from_orderable((uint32_t)(k >> 32));
            score = 1.f / (1.f + expf(-logit));
            label = flat % p.C;
            qidx = flat / p.C;
            const float* q = p.box + ((int64_t)b * p.Q + qidx) * p.code;
            bx[0] = q[0]; bx[1] = q[1]; bx[2] = q[4];
            bx[3] = expf(q[2]); bx[4] = expf(q[3]); bx[5] = expf(q[5]);
            bx[6] = atan2f(q[6], q[7]);
            if (p.code > 8) { bx[7] = q[8]; bx[8] = q[9]; }
            keep = bx[0] >= p.range[0] && bx[1] >= p.range[1] && bx[2] >= p.range[2] &&
                   bx[0] <= p.range[3] && bx[1] <= p.range[4] && bx[2] <= p.range[5];
            if (p.has_thr) keep = keep && (score > p.thr);
            if (p.bottom_center) bx[2] = bx[2] - bx[5] * 0.5f;
        }
        for (int i = 0; i < nb; ++i) p.out_boxes[o * nb + i] = bx[i];
        p.out_scores[o] = score;
        p.out_labels[o] = label;
        p.out_index[o] = qidx;
        p.out_mask[o] = keep ? 1 : 0;
        if (keep) atomicAdd(&s_count, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) p.out_count[b] = s_count;
}

}  // namespace
}  // namespace petr

using namespace petr;

extern "C" int petr_nms_free_decode(const float* cls_scores, const float* bbox_preds, int B, int Q, int C, int code,
                                    int max_num, const float* h_post_center_range, int has_score_threshold,
                                    float score_threshold, int bottom_center, float* out_boxes, float* out_scores,
                                    int64_t* out_labels, int64_t* out_index, uint8_t* out_mask, int32_t* out_count,
                                    void* stream) {
    PETR_REQUIRE(cls_scores && bbox_preds && h_post_center_range && out_boxes && out_scores && out_labels && out_index &&
                 out_mask && out_count, PETR_E_BADARG, "petr_nms_free_decode: null pointer");
    PETR_REQUIRE(B > 0 && Q > 0 && C > 0 && max_num > 0, PETR_E_BADARG, "petr_nms_free_decode: bad dimension");
    PETR_REQUIRE(code == 8 || code == 10, PETR_E_SHAPE, "petr_nms_free_decode: code_size must be 8 or 10, got %d", code);
    int nsort = 2;
    while (nsort < Q * C) nsort <<= 1;
    if (nsort < max_num) { while (nsort < max_num) nsort <<= 1; }
    PETR_REQUIRE(nsort <= 16384, PETR_E_SHAPE, "petr_nms_free_decode: num_query * num_classes = %d exceeds 16384", Q * C);
    DecodeParams p{};
    p.cls = cls_scores; p.box = bbox_preds; p.Q = Q; p.C = C; p.code = code; p.max_num = max_num; p.nsort = nsort;
    p.bottom_center = bottom_center; p.has_thr = has_score_threshold; p.thr = score_threshold;
    for (int i = 0; i < 6; ++i) p.range[i] = h_post_center_range[i];
    p.out_boxes = out_boxes; p.out_scores = out_scores; p.out_labels = out_labels; p.out_index = out_index;
    p.out_mask = out_mask; p.out_count = out_count;
    const int smem = nsort * 8;
    cudaError_t e = cudaFuncSetAttribute(nms_free_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail((int)e, "petr_nms_free_decode: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    nms_free_decode_kernel<<<B, DEC_THREADS, smem, (cudaStream_t)stream>>>(p);
    return after_launch("petr_nms_free_decode");
}
