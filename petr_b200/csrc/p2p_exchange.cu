// Key-split over camera views without a library collective: the (O, m, l) partials of a rank's
// flash attention are PUSHED into every peer's gather buffer with plain stores over NVLink (peer-mapped symmetric
// memory), a per-(parity, source rank) flag is released at system scope by the last CTA of the push, and the merge
// kernel acquires the flags of all source ranks before it reads.  No NCCL call on the data path, so the whole head
// (with the exchange) is CUDA-graph capturable and replays as one graph per rank.
//
//   buffers of rank y (one symmetric allocation):  gathered[2 parities][world][slot_floats] fp32, flags[2][world] u32
//   layer l of forward e:  parity = l & 1,  target = e * seqs_per_epoch + l + 1  (e read from device memory at run time,
//   incremented by petr_p2p_epoch_inc once per forward -- inside the captured graph)
//   write-after-read safety: rank r pushes layer l+2 (same parity) only after it merged layer l+1, which waited for
//   every peer's layer-(l+1) push, which each peer issued after its own layer-l merge (stream order).
#include "common.cuh"

namespace petr {
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct PushParams {
    const float4* src; void* const* peer_bufs; void* const* peer_flags; int world, rank;
    int64_t slot_floats; int parity; const uint32_t* epoch; int seq, seqs_per_epoch; uint32_t* counters;
};

// grid (ctas_per_peer, world).  CTA (x, y): slice x of this rank's partial -> peer y's slot [parity][rank]; the last CTA
// of a peer (per-peer counter, atomicInc wraps to 0 for the next launch / graph replay) releases the flag.
__global__ void __launch_bounds__(256)
p2p_push_kernel(PushParams p) {
    const int y = blockIdx.y;
    const int64_t n4 = p.slot_floats / 4;
    if (y != p.rank) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.peer_bufs[y]) +
                                                ((int64_t)p.parity * p.world + p.rank) * p.slot_floats);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
            dst[i] = p.src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicInc(p.counters + y, gridDim.x - 1);
        if (done == gridDim.x - 1) {
            __threadfence_system();
            const uint32_t target = (*p.epoch) * (uint32_t)p.seqs_per_epoch + (uint32_t)p.seq + 1u;
            st_release_sys(reinterpret_cast<uint32_t*>(p.peer_flags[y]) + p.parity * p.world + p.rank, target);
        }
    }
}

__global__ void epoch_inc_kernel(uint32_t* epoch) { *epoch += 1u; }

// merge of `nparts` gathered partials (layout of petr_attention_merge) after acquiring the flags of all source ranks
__global__ void __launch_bounds__(256)
attention_merge_p2p_kernel(const float* __restrict__ part, int nparts, int nflags, const uint32_t* flags, const uint32_t* epoch,
                           int seq, int seqs_per_epoch, uint32_t* err, float* __restrict__ O, int ldo,
                           __nv_bfloat16* __restrict__ O_split, int B, int Lq, int heads) {
    if (threadIdx.x < nflags) {
        const uint32_t target = (*epoch) * (uint32_t)seqs_per_epoch + (uint32_t)seq + 1u;
        const long long t0 = clock64();
        // monotonic flags: >= target (a peer may already be one layer ahead on the other parity, never on this one)
        while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - target) < 0) {
            if (clock64() - t0 > 4000000000ll) { atomicExch(err, 1u + threadIdx.x); break; }     // ~2 s: report, do not hang
            __nanosleep(200);
        }
    }
    __syncthreads();
    const int64_t rows = (int64_t)B * heads * Lq;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * 8) return;
    const int64_t r = idx >> 3;
    const int d4 = (int)(idx & 7) * 4;
    const int64_t part_stride = rows * 34;
    float m = -INFINITY;
    for (int pz = 0; pz < nparts; ++pz) m = fmaxf(m, __ldcg(part + pz * part_stride + rows * 32 + r * 2));
    float den = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pz = 0; pz < nparts; ++pz) {
        const float* base = part + pz * part_stride;
        const float mp = __ldcg(base + rows * 32 + r * 2), lp = __ldcg(base + rows * 32 + r * 2 + 1);
        const float w = (mp == -INFINITY) ? 0.f : exp2f(mp - m);
        den += w * lp;
        const float4 o = __ldcg(reinterpret_cast<const float4*>(base + r * 32 + d4));
        acc.x += w * o.x; acc.y += w * o.y; acc.z += w * o.z; acc.w += w * o.w;
    }
    const float inv = 1.f / den;
    const int q = (int)(r % Lq);
    const int bh = (int)(r / Lq);
    const int b = bh / heads, h = bh % heads;
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    if (O) *reinterpret_cast<float4*>(O + ((int64_t)b * Lq + q) * ldo + h * 32 + d4) = acc;
    if (O_split) {
        const int E = heads * 32;
        __nv_bfloat16* sp = O_split + ((int64_t)b * Lq + q) * (2 * E) + h * 32 + d4;
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y), h1 = __floats2bfloat162_rn(acc.z, acc.w);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(acc.x - __low2float(h0), acc.y - __high2float(h0));
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(acc.z - __low2float(h1), acc.w - __high2float(h1));
        *reinterpret_cast<uint2*>(sp) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(sp + E) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
}

}  // namespace
}  // namespace petr

using namespace petr;

extern "C" int petr_p2p_push_partial(const float* src, void* const* d_peer_bufs, void* const* d_peer_flags, int world,
                                     int rank, int64_t slot_floats, int parity, const uint32_t* d_epoch, int seq,
                                     int seqs_per_epoch, uint32_t* d_counters, void* stream) {
    PETR_REQUIRE(src && d_peer_bufs && d_peer_flags && d_epoch && d_counters, PETR_E_BADARG, "petr_p2p_push_partial: null pointer");
    PETR_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && (parity == 0 || parity == 1) && seq >= 0 &&
                 seq < seqs_per_epoch, PETR_E_BADARG, "petr_p2p_push_partial: bad world / rank / parity / seq");
    PETR_REQUIRE(slot_floats > 0 && (slot_floats % 4) == 0 && aligned16(src), PETR_E_ALIGN,
                 "petr_p2p_push_partial: slot must be a multiple of 4 floats, 16-byte aligned");
    PushParams p{reinterpret_cast<const float4*>(src), d_peer_bufs, d_peer_flags, world, rank, slot_floats, parity, d_epoch,
                 seq, seqs_per_epoch, d_counters};
    int ctas = (int)((slot_floats / 4 + 256 * 8 - 1) / (256 * 8));
    if (ctas > 64) ctas = 64;
    if (ctas < 1) ctas = 1;
    p2p_push_kernel<<<dim3(ctas, world), 256, 0, (cudaStream_t)stream>>>(p);
    return after_launch("petr_p2p_push_partial");
}

extern "C" int petr_p2p_epoch_inc(uint32_t* d_epoch, void* stream) {
    PETR_REQUIRE(d_epoch, PETR_E_BADARG, "petr_p2p_epoch_inc: null pointer");
    epoch_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_epoch);
    return after_launch("petr_p2p_epoch_inc");
}

extern "C" int petr_attention_merge_p2p(const float* gathered, int nparts, int world, const uint32_t* d_flags, const uint32_t* d_epoch,
                                        int seq, int seqs_per_epoch, uint32_t* d_err, float* O, int ldo, void* O_split,
                                        int B, int Lq, int heads, void* stream) {
    PETR_REQUIRE(gathered && d_flags && d_epoch && d_err && (O || O_split), PETR_E_BADARG, "petr_attention_merge_p2p: null pointer");
    PETR_REQUIRE(world >= 1 && world <= 64 && nparts >= world && nparts % world == 0 && B > 0 && Lq > 0 && heads > 0, PETR_E_BADARG,
                 "petr_attention_merge_p2p: bad dimension (nparts must be a multiple of world)");
    const int64_t n = (int64_t)B * heads * Lq * 8;
    attention_merge_p2p_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        gathered, nparts, world, d_flags, d_epoch, seq, seqs_per_epoch, d_err, O, ldo, (__nv_bfloat16*)O_split, B, Lq, heads);
    return after_launch("petr_attention_merge_p2p");
}
