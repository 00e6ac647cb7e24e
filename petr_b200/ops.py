"""Tensor-level wrappers over the C ABI (one function per entry point of include/petr_b200.h).

PyTorch is used for device memory and streams only; every computation is a kernel of
libpetr_b200.so launched on torch's current CUDA stream.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import call, host_floats, ptr, stream_ptr

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_SIGMOID_GATE = 0, 1, 2
A_ROWMAJOR, A_NCHW = 0, 1


class _Timed:
    """Labels the C-ABI calls made inside the block for _lib.TIMERS (bench.py)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        _lib.LABEL = self.name

    def __exit__(self, *a):
        _lib.LABEL = None


def _cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"petr_b200: {name} must be a CUDA tensor (no CPU fallback)")


def _dt(dtype):
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"petr_b200: unsupported dtype {dtype}")


def key_padding_mask(img_hw: torch.Tensor, pad_h: int, pad_w: int, H: int, W: int) -> torch.Tensor:
    """img_hw int32 (B,N,2) on device -> uint8 (B,N,H,W); petr_head.py:383-394."""
    _cuda(img_hw, "img_hw")
    B, N, _ = img_hw.shape
    out = torch.empty((B, N, H, W), dtype=torch.uint8, device=img_hw.device)
    call("petr_key_padding_mask", ptr(img_hw), B, N, int(pad_h), int(pad_w), H, W, ptr(out),
         stream_ptr(img_hw.device))
    return out


def frustum_features(img2lidar: torch.Tensor, masks: Optional[torch.Tensor], H: int, W: int, D: int,
                     pad_h: int, pad_w: int, lid: bool, depth_start: float, bin_size: float,
                     position_range, dtype=torch.float32, want_mask=True, want_coords=False):
    """petr_head.py:286-331.  Returns (feat (T,3D), coords_mask uint8 (B,N,H,W) | None,
    coords3d (B,N,H,W,D,3) | None)."""
    _cuda(img2lidar, "img2lidar")
    B, N = img2lidar.shape[:2]
    dev = img2lidar.device
    T = B * N * H * W
    feat = torch.empty((T, 3 * D), dtype=dtype, device=dev)
    cmask = torch.empty((B, N, H, W), dtype=torch.uint8, device=dev) if want_mask else None
    c3 = torch.empty((B, N, H, W, D, 3), dtype=torch.float32, device=dev) if want_coords else None
    pr = [float(v) for v in position_range]
    lo = host_floats(pr[:3])
    span = host_floats([pr[3] - pr[0], pr[4] - pr[1], pr[5] - pr[2]])
    call("petr_frustum_features", ptr(img2lidar), ptr(masks), B, N, H, W, D, int(pad_h), int(pad_w),
         int(bool(lid)), float(depth_start), float(bin_size), lo, span, ptr(feat), _dt(dtype),
         ptr(cmask), ptr(c3), stream_ptr(dev))
    return feat, cmask, c3


def sine3d_features(mask: torch.Tensor, dim_t: torch.Tensor, num_feats: int, normalize: bool,
                    scale: float, eps: float, offset: float, dtype=torch.float32) -> torch.Tensor:
    """positional_encoding.py:58-100.  mask uint8 (B,N,H,W) -> (T, 3*num_feats)."""
    _cuda(mask, "mask")
    B, N, H, W = mask.shape
    out = torch.empty((B * N * H * W, 3 * num_feats), dtype=dtype, device=mask.device)
    call("petr_sine3d_features", ptr(mask), ptr(dim_t), B, N, H, W, num_feats, int(bool(normalize)),
         float(scale), float(eps), float(offset), ptr(out), _dt(dtype), stream_ptr(mask.device))
    return out


def pos2posemb3d(ref: torch.Tensor, dim_t: torch.Tensor) -> torch.Tensor:
    """petr_head.py:31-43.  ref (Q,3) fp32 -> (Q, 3*F)."""
    _cuda(ref, "ref")
    Q, Fn = ref.shape[0], dim_t.numel()
    out = torch.empty((Q, 3 * Fn), dtype=torch.float32, device=ref.device)
    call("petr_pos2posemb3d", ptr(ref), ptr(dim_t), Q, Fn, ptr(out), stream_ptr(ref.device))
    return out


def gemm_f32(A: torch.Tensor, W: torch.Tensor, bias=None, *, A2=None, residual=None, gate=None,
             act=ACT_NONE, out=None, out_bf16=None, nchw_hw: int = 0, M: Optional[int] = None,
             want_f32: bool = True) -> torch.Tensor:
    """C = epi((A + A2[m % rows]) @ W.T + bias) + residual, fp32 accumulate.

    A: (M,K) with unit inner stride, or (when nchw_hw > 0) a contiguous (G,K,h,w) map read as
    M = G*h*w rows.  W: (N,K) contiguous."""
    _cuda(A, "A")
    N, K = W.shape[0], W[0].numel()
    assert W.is_contiguous()
    if nchw_hw > 0:
        assert A.is_contiguous()
        M = A.numel() // K
        lda, layout = K, A_NCHW
    else:
        assert A.dim() == 2 and A.stride(1) == 1 and A.shape[1] == K
        M = A.shape[0] if M is None else M
        lda, layout = A.stride(0), A_ROWMAJOR
    a2_rows = 0
    if A2 is not None:
        assert A2.dim() == 2 and A2.stride(1) == 1 and A2.stride(0) == lda and A2.shape[1] == K
        a2_rows = A2.shape[0]
    if out is None and want_f32:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    ref_t = out if out is not None else out_bf16
    assert ref_t.stride(-1) == 1
    ldc = ref_t.stride(0)
    for t in (residual, gate, out_bf16):
        if t is not None:
            assert t.stride(0) == ldc and t.stride(1) == 1
    call("petr_gemm_f32", ptr(A), lda, layout, int(nchw_hw), ptr(A2), a2_rows, ptr(W), ptr(bias),
         ptr(residual), ptr(gate), act, ptr(out), ptr(out_bf16), ldc, M, N, K, stream_ptr(A.device))
    return out if out is not None else out_bf16


def attention_f32(Q, K, V, key_mask, B: int, Lq: int, Lk: int, heads: int, out=None,
                  out_split=None, want_f32: bool = True, attn_bias=None, lse=None, drop_p: float = 0.0,
                  seed: int = 0) -> torch.Tensor:
    """softmax(Q K^T / sqrt(d) + mask) V; Q rows b*Lq+q, K/V rows b*Lk+k, head_dim 32.
    out_split: optional bf16 (B*Lq, 2E) [hi | lo] copy of the result.  attn_bias: additive fp32 (Lq, Lk) or
    (B*heads, Lq, Lk).  lse / drop_p / seed: training (see petr_attention_f32_bwd)."""
    _cuda(Q, "Q")
    d = 32
    E = heads * d
    if out is None and want_f32:
        out = torch.empty((B * Lq, E), dtype=torch.float32, device=Q.device)
    per_head = 0
    if attn_bias is not None:
        assert attn_bias.dtype == torch.float32 and attn_bias.is_contiguous()
        assert tuple(attn_bias.shape) in ((Lq, Lk), (B * heads, Lq, Lk)), attn_bias.shape
        per_head = int(attn_bias.dim() == 3)
    # one thread per query row, 64 rows per CTA, 144 registers: seven CTAs per SM.  Split the key axis until the CTAs
    # that have work (a last row block less than half full ends at once) make two full waves; each part >= 8 key tiles
    busy = max(1, (Lq // 64 + (1 if Lq % 64 >= 32 else 0)) * heads * B)
    ksplit = max(1, min((2 * 7 * 148) // busy, ((Lk + 31) // 32) // 8, 65535 // max(B, 1)))
    partial = None
    if ksplit > 1:
        partial = torch.empty((ksplit * attention_partial_floats(B, Lq, heads),), dtype=torch.float32, device=Q.device)
    with _Timed("attention_f32"):
        call("petr_attention_f32", ptr(Q), Q.stride(0), ptr(K), K.stride(0), ptr(V), V.stride(0),
             ptr(key_mask), ptr(attn_bias), per_head, ptr(out), out.stride(0) if out is not None else 0,
             ptr(out_split), ptr(lse), float(drop_p), int(seed), ptr(partial), ksplit, B, Lq, Lk, heads, d,
             1.0 / math.sqrt(d), stream_ptr(Q.device))
    return out if out is not None else out_split


def layernorm_multi(X, gamma, beta, *, eps=1e-5, qpos=None, Y=None, Y_split=None, Ypos_split=None,
                    gamma2=None, beta2=None, Z=None, Z_split=None, relu=False, nparts: int = 1):
    """Fused LayerNorm with optional fp32 / split-bf16 / (+qpos) / second-LN outputs (E = 256).
    nparts > 1: X is (nparts, rows, 256), the split-K parts of the preceding GEMM, summed on load."""
    _cuda(X, "X")
    assert X.is_contiguous() and X.shape[-1] == 256
    rows = X.numel() // 256 // nparts
    call("petr_layernorm_multi_f32", ptr(X), ptr(gamma), ptr(beta), float(eps), ptr(qpos),
         qpos.shape[0] if qpos is not None else 0, ptr(Y), ptr(Y_split), ptr(Ypos_split),
         ptr(gamma2), ptr(beta2), ptr(Z), ptr(Z_split), rows, 256, int(relu), int(nparts), rows * 256,
         stream_ptr(X.device))


def split_rows(X, add=None, relu=False, out=None) -> torch.Tensor:
    """fp32 (R,E) -> bf16 (R,2E) [hi | lo] of relu?(X + add[r % add_rows])."""
    _cuda(X, "X")
    R, E = X.shape
    if out is None:
        out = torch.empty((R, 2 * E), dtype=torch.bfloat16, device=X.device)
    call("petr_split_rows_bf16", ptr(X), X.stride(0), ptr(add), add.shape[0] if add is not None else 0,
         ptr(out), R, E, int(relu), stream_ptr(X.device))
    return out


def nchw_to_tokens_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 | bf16 (G,C,h,w) contiguous -> bf16 (G*h*w, C)."""
    _cuda(x, "x")
    assert x.is_contiguous() and x.dim() == 4
    G, C, h, w = x.shape
    out = torch.empty((G * h * w, C), dtype=torch.bfloat16, device=x.device)
    call("petr_nchw_to_tokens_bf16", ptr(x), _dt(x.dtype), ptr(out), G, C, h * w, stream_ptr(x.device))
    return out


def layernorm(X, gamma, beta, eps=1e-5, relu=False, nan_to_num=False, out=None) -> torch.Tensor:
    _cuda(X, "X")
    assert X.is_contiguous()
    E = X.shape[-1]
    rows = X.numel() // E
    if out is None:
        out = torch.empty_like(X)
    call("petr_layernorm_f32", ptr(X), ptr(gamma), ptr(beta), ptr(out), rows, E, float(eps),
         int(relu), int(nan_to_num), stream_ptr(X.device))
    return out


def decode_boxes(reg: torch.Tensor, ref: torch.Tensor, pc_range, time_div, L: int, B: int, Q: int):
    """In place on reg (L*B*Q, code_size); petr_head.py:440-460."""
    _cuda(reg, "reg")
    assert reg.is_contiguous()
    call("petr_decode_boxes", ptr(reg), ptr(ref), host_floats(list(pc_range)), ptr(time_div),
         L, B, Q, reg.shape[-1], stream_ptr(reg.device))
    return reg


def launch_count() -> int:
    return _lib.launch_count()


def reset_launch_count():
    _lib.reset_launch_count()


# ---- tcgen05 tier ---------------------------------------------------------------------------
OUT_F32, OUT_BF16, OUT_HEADSPLIT = 0, 1, 2


def x3_weight(W: torch.Tensor) -> torch.Tensor:
    """fp32 (N,K) -> bf16 (N,3K) = [W_hi | W_hi | W_lo], the W operand of a bf16x3 GEMM."""
    hi = W.to(torch.bfloat16)
    lo = (W - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], 1).contiguous()


def x3_weight_kernel(W: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """x3_weight(W) (or of W.T) in one launch of the library: fp32 (N,K) -> bf16 (N,3K) | (K,3N)."""
    _cuda(W, "W")
    assert W.dim() == 2 and W.is_contiguous() and W.dtype == torch.float32
    N, K = W.shape
    out = torch.empty((K, 3 * N) if transpose else (N, 3 * K), dtype=torch.bfloat16, device=W.device)
    call("petr_x3_weight_bf16", ptr(W), N, K, int(bool(transpose)), ptr(out), stream_ptr(W.device))
    return out


def split_transpose(X: torch.Tensor, add, Mpad: int, w3: bool) -> torch.Tensor:
    """fp32 (M,C) (+ add[m % rows]) -> bf16 (C, 2*Mpad) [hi | lo] or (C, 3*Mpad) [hi | hi | lo] of the transpose, zero
    beyond column M: the operands of the weight-gradient GEMM."""
    _cuda(X, "X")
    assert X.dim() == 2 and X.stride(1) == 1 and X.dtype == torch.float32
    M, C = X.shape
    out = torch.empty((C, (3 if w3 else 2) * Mpad), dtype=torch.bfloat16, device=X.device)
    if add is not None:
        assert add.is_contiguous() and add.shape[1] == C
    call("petr_split_transpose_bf16", ptr(X), X.stride(0), ptr(add), add.shape[0] if add is not None else 0, M, C,
         int(Mpad), int(bool(w3)), ptr(out), stream_ptr(X.device))
    return out


OUT_SPLIT = 3


def gemm_bf16_tc(A: torch.Tensor, W: torch.Tensor, bias=None, *, residual=None, gate=None, act=ACT_NONE,
                 alpha: float = 1.0, out=None, out_dtype=torch.float32, headsplit_Lk: int = 0,
                 split_out: bool = False, out_f32=None, x3: bool = False, headsplit_pair: bool = False,
                 residual_period: int = 0, ksplit: int = 1, name: str = "gemm_bf16_tc"):
    """C = alpha * act(A @ W.T + bias) (+ residual) on tcgen05; A (M,Ka) bf16, W (N,K) bf16.
    x3: A is a split (M, 2*K0) [hi | lo] activation and W an x3_weight (N, 3*K0).
    headsplit_Lk > 0: bf16 output in the [N/256][B][8][Lk][32] Q/K/V layout.
    split_out: bf16 (M, 2N) [hi | lo] output (+ optional fp32 copy out_f32)."""
    _cuda(A, "A")
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    assert A.dim() == 2 and A.stride(1) == 1 and W.dim() == 2 and W.stride(1) == 1
    M = A.shape[0]
    N, K = W.shape
    wrap = 0
    if x3:
        assert 3 * A.shape[1] == 2 * K, (A.shape, W.shape)
        wrap = A.shape[1]
    else:
        assert A.shape[1] == K
    C2 = None
    if residual_period > 0:
        # out_mode 5: bf16 result without the residual + fp32 result with residual[m % period]
        assert residual is not None and out_f32 is not None and residual.shape[0] == residual_period
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16, device=A.device)
        with _Timed(name):
            call("petr_gemm_bf16_tc", ptr(A), A.stride(0), wrap, ptr(W), W.stride(0), ptr(bias), ptr(residual), None,
                 act, float(alpha), ptr(out), ptr(out_f32), out_f32.stride(0), 5, M, N, K, int(residual_period), 1,
                 stream_ptr(A.device))
        return out
    if headsplit_Lk > 0 and headsplit_pair:
        # (hi, lo) pair stacked on a leading axis: out[0] = hi, out[1] = lo, each [N/256][B][8][L][32]
        B = M // headsplit_Lk
        if out is None:
            out = torch.empty((2, N // 256, B, 8, headsplit_Lk, 32), dtype=torch.bfloat16, device=A.device)
        C2 = out[1]
        mode, ldc = 4, 0
    elif headsplit_Lk > 0:
        B = M // headsplit_Lk
        if out is None:
            out = torch.empty((N // 256, B, 8, headsplit_Lk, 32), dtype=torch.bfloat16, device=A.device)
        mode, ldc = OUT_HEADSPLIT, 0
    elif split_out:
        if out is None:
            out = torch.empty((M, 2 * N), dtype=torch.bfloat16, device=A.device)
        mode = OUT_SPLIT
        C2 = out_f32
        ldc = C2.stride(0) if C2 is not None else (residual.stride(0) if residual is not None else N)
    else:
        if out is None:
            shape = (M, N) if ksplit <= 1 else (ksplit, M, N)         # split-K: one partial product per part
            out = torch.empty(shape, dtype=out_dtype, device=A.device)
        mode = OUT_F32 if out.dtype == torch.float32 else OUT_BF16
        ldc = out.stride(-2)
        assert ksplit <= 1 or (mode == OUT_F32 and gate is None and out.is_contiguous())
    for t in (residual, gate):
        if t is not None:
            assert t.dtype == torch.float32 and t.stride(0) == ldc and t.stride(1) == 1
    with _Timed(name):
        call("petr_gemm_bf16_tc", ptr(A), A.stride(0), wrap, ptr(W), W.stride(0), ptr(bias), ptr(residual), ptr(gate),
             act, float(alpha), ptr(out), ptr(C2), ldc, mode, M, N, K, int(headsplit_Lk), int(max(1, ksplit)),
             stream_ptr(A.device))
    return out


def position_mlp_fused(img2lidar, H: int, W: int, D: int, pad_h: int, pad_w: int, lid: bool,
                       depth_start: float, bin_size: float, position_range, W1, b1, W2, b2,
                       residual=None, out=None, residual2=None, out_bf16=None, want_f32: bool = True):
    """Fused position embedding (petr_head.py:286-334) on tcgen05: img2lidar (B,N,4,4) fp32 ->
    fp32 tokens (T, 256) = MLP(inverse_sigmoid(frustum coords)) (+ residual).  W1 (1024,192) /
    W2 (256,1024) bf16."""
    _cuda(img2lidar, "img2lidar")
    B, N = img2lidar.shape[:2]
    T = B * N * H * W
    hidden, embed = W1.shape[0], W2.shape[0]
    assert W1.dtype == torch.bfloat16 and W2.dtype == torch.bfloat16 and W1.is_contiguous() and W2.is_contiguous()
    if out is None and want_f32:
        out = torch.empty((T, embed), dtype=torch.float32, device=img2lidar.device)
    pr = [float(v) for v in position_range]
    lo = host_floats(pr[:3])
    span = host_floats([pr[3] - pr[0], pr[4] - pr[1], pr[5] - pr[2]])
    with _Timed("position_mlp_fused"):
        call("petr_position_mlp_fused", ptr(img2lidar), B, N, H, W, D, int(pad_h), int(pad_w), int(bool(lid)),
             float(depth_start), float(bin_size), lo, span, ptr(W1), ptr(b1), ptr(W2), ptr(b2), hidden, embed,
             ptr(residual), ptr(residual2), residual2.shape[0] if residual2 is not None else 0, ptr(out),
             ptr(out_bf16), stream_ptr(img2lidar.device))
    return out if out is not None else out_bf16


def add_rows_bf16(X, add, out=None) -> torch.Tensor:
    """bf16(X[r] + add[r % add_rows]) for fp32 (R,E), (add_rows,E)."""
    _cuda(X, "X")
    assert X.is_contiguous() and add.is_contiguous()
    R, E = X.shape
    if out is None:
        out = torch.empty((R, E), dtype=torch.bfloat16, device=X.device)
    call("petr_add_rows_bf16", ptr(X), ptr(add), add.shape[0], ptr(out), R, E, stream_ptr(X.device))
    return out


def pack_q_bf16(Q: torch.Tensor, B: int, Lq: int, heads: int, out=None) -> torch.Tensor:
    """fp32 (B*Lq, heads*32) -> bf16 (B, heads, Lq, 32) scaled by 32^-1/2 * log2(e)."""
    _cuda(Q, "Q")
    if out is None:
        out = torch.empty((B, heads, Lq, 32), dtype=torch.bfloat16, device=Q.device)
    scale = (1.0 / math.sqrt(32.0)) * 1.4426950408889634
    call("petr_pack_q_bf16", ptr(Q), Q.stride(0), ptr(out), B, Lq, heads, scale, stream_ptr(Q.device))
    return out


def attention_partial_floats(B: int, Lq: int, heads: int) -> int:
    return int(_lib.lib().petr_attention_partial_floats(B, Lq, heads))


def cross_attention_workspace(B: int, Lq: int, Lk: int, heads: int, nsplit: int, device) -> torch.Tensor:
    """Scratch of petr_cross_attention_bf16's fast pass (one flag per CTA)."""
    n = int(_lib.lib().petr_cross_attention_workspace_bytes(B, Lq, Lk, heads, nsplit))
    return torch.empty((max(n, 4),), dtype=torch.uint8, device=device)


def cross_attention_bf16(Qs, Kh, Vh, key_mask, B: int, Lq: int, Lk: int, heads: int, nsplit: int = 1,
                         out=None, partial=None, write_partial: bool = False, out_split=None,
                         want_f32: bool = True, workspace=None, exact_only: bool = False):
    """Flash cross-attention on tcgen05.  Returns O fp32 (B*Lq, heads*32), or the partial
    buffer when write_partial.  workspace: cross_attention_workspace(...) (allocated here when
    None); exact_only skips the max-free fast pass."""
    _cuda(Qs, "Qs")
    nsplit = max(1, min(int(nsplit), (Lk + 127) // 128))
    if workspace is None and not exact_only:
        workspace = cross_attention_workspace(B, Lq, Lk, heads, nsplit, Qs.device)
    if exact_only:
        workspace = None
    if (nsplit > 1 or write_partial) and partial is None:
        partial = torch.empty((nsplit * attention_partial_floats(B, Lq, heads),), dtype=torch.float32,
                              device=Qs.device)
    if not write_partial and out is None and want_f32:
        out = torch.empty((B * Lq, heads * 32), dtype=torch.float32, device=Qs.device)
    with _Timed("cross_attention_bf16"):
        call("petr_cross_attention_bf16", ptr(Qs), ptr(Kh), ptr(Vh), ptr(key_mask), ptr(out),
             out.stride(0) if out is not None else 0, ptr(out_split), ptr(partial), ptr(workspace),
             B, Lq, Lk, heads, nsplit, int(write_partial), stream_ptr(Qs.device))
    if write_partial:
        return partial
    return out if out is not None else out_split


def attention_x3_tc(Qp, Kp, Vp, key_mask, B: int, Lq: int, Lk: int, heads: int, nsplit: int = 1,
                    out=None, out_split=None, partial=None, write_partial: bool = False,
                    want_f32: bool = True, workspace=None, exact_only: bool = False):
    """fp32-class attention on tcgen05.  Qp/Kp/Vp: bf16 (2, B, heads, L, 32) = (hi, lo) pairs.  workspace: flag
    scratch of the fast pass (allocated here when None); exact_only skips the max-free fast pass."""
    _cuda(Qp, "Qp")
    nsplit = max(1, min(int(nsplit), (Lk + 127) // 128))
    if exact_only:
        workspace = None
    elif workspace is None:
        n = int(_lib.lib().petr_attention_x3_workspace_bytes(B, Lq, Lk, heads, nsplit))
        workspace = torch.empty((max(n, 4),), dtype=torch.uint8, device=Qp.device)
    if (nsplit > 1 or write_partial) and partial is None:
        partial = torch.empty((nsplit * attention_partial_floats(B, Lq, heads),), dtype=torch.float32,
                              device=Qp.device)
    if not write_partial and out is None and want_f32:
        out = torch.empty((B * Lq, heads * 32), dtype=torch.float32, device=Qp.device)
    with _Timed("attention_x3_tc"):
        call("petr_attention_x3_tc", ptr(Qp[0]), ptr(Qp[1]), ptr(Kp[0]), ptr(Kp[1]), ptr(Vp[0]), ptr(Vp[1]),
             ptr(key_mask), ptr(out), out.stride(0) if out is not None else 0, ptr(out_split), ptr(partial),
             ptr(workspace), B, Lq, Lk, heads, 1.0 / math.sqrt(32.0), nsplit, int(write_partial), stream_ptr(Qp.device))
    if write_partial:
        return partial
    return out if out is not None else out_split


def attention_merge(partial: torch.Tensor, nparts: int, B: int, Lq: int, heads: int, out=None,
                    out_split=None, want_f32: bool = True):
    _cuda(partial, "partial")
    if out is None and want_f32:
        out = torch.empty((B * Lq, heads * 32), dtype=torch.float32, device=partial.device)
    call("petr_attention_merge", ptr(partial), nparts, ptr(out), out.stride(0) if out is not None else 0,
         ptr(out_split), None, B, Lq, heads, stream_ptr(partial.device))
    return out if out is not None else out_split


def nms_free_decode(cls_scores: torch.Tensor, bbox_preds: torch.Tensor, max_num: int, post_center_range,
                    score_threshold=None, bottom_center: bool = False):
    """NMSFreeCoder.decode_single for a batch (nms_free_coder.py:48-101), one launch.  cls_scores (B,Q,C) logits,
    bbox_preds (B,Q,code) fp32 CUDA.  Returns padded (boxes (B,max_num,9|7), scores, labels int64, query index
    int64, mask bool, count int32); rows are sorted by descending score."""
    _cuda(cls_scores, "cls_scores")
    cls_scores = cls_scores.float().contiguous()
    bbox_preds = bbox_preds.float().contiguous()
    B, Q, C = cls_scores.shape
    code = bbox_preds.shape[-1]
    dev = cls_scores.device
    nb = code - 1 if code > 8 else 7
    boxes = torch.empty((B, max_num, nb), dtype=torch.float32, device=dev)
    scores = torch.empty((B, max_num), dtype=torch.float32, device=dev)
    labels = torch.empty((B, max_num), dtype=torch.int64, device=dev)
    index = torch.empty((B, max_num), dtype=torch.int64, device=dev)
    mask = torch.empty((B, max_num), dtype=torch.uint8, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    thr = score_threshold
    call("petr_nms_free_decode", ptr(cls_scores), ptr(bbox_preds), B, Q, C, code, int(max_num),
         host_floats([float(v) for v in post_center_range]), int(thr is not None), float(thr or 0.0),
         int(bool(bottom_center)), ptr(boxes), ptr(scores), ptr(labels), ptr(index), ptr(mask), ptr(count),
         stream_ptr(dev))
    return boxes, scores, labels, index, mask.bool(), count


# ---- either side of the main path (2-D positional encodings, CPFPN tail, NormedLinear) -----------------------
def sine2d_features(mask: torch.Tensor, dim_t: torch.Tensor, num_feats: int, normalize: bool, scale: float,
                    eps: float, offset: float, dtype=torch.float32) -> torch.Tensor:
    """mmdet SinePositionalEncoding.  mask uint8 (G,H,W) -> (G*H*W, 2*num_feats)."""
    _cuda(mask, "mask")
    G, H, W = mask.shape
    out = torch.empty((G * H * W, 2 * num_feats), dtype=dtype, device=mask.device)
    call("petr_sine2d_features", ptr(mask), ptr(dim_t), G, H, W, num_feats, int(bool(normalize)), float(scale),
         float(eps), float(offset), ptr(out), _dt(dtype), stream_ptr(mask.device))
    return out


def learned_pos_features(row_embed: torch.Tensor, col_embed: torch.Tensor, G: int, H: int, W: int) -> torch.Tensor:
    """positional_encoding.py:137-160 -> (G*H*W, 2F) fp32 tokens."""
    _cuda(row_embed, "row_embed")
    Fn = row_embed.shape[1]
    out = torch.empty((G * H * W, 2 * Fn), dtype=torch.float32, device=row_embed.device)
    call("petr_learned_pos_features", ptr(row_embed), row_embed.shape[0], ptr(col_embed), col_embed.shape[0],
         G, H, W, Fn, ptr(out), stream_ptr(row_embed.device))
    return out


def tokens_to_nchw(x: torch.Tensor, G: int, hw: int) -> torch.Tensor:
    """fp32 (G*hw, C) -> (G, C, hw)."""
    _cuda(x, "x")
    assert x.is_contiguous() and x.dtype == torch.float32
    C = x.shape[-1]
    out = torch.empty((G, C, hw), dtype=torch.float32, device=x.device)
    call("petr_tokens_to_nchw_f32", ptr(x), ptr(out), G, C, hw, stream_ptr(x.device))
    return out


def upsample_add_tokens(dst: torch.Tensor, src: torch.Tensor, G: int, H: int, W: int, h: int, w: int):
    """dst (G*H*W, C) += nearest-upsampled src (G*h*w, C), in place."""
    _cuda(dst, "dst")
    assert dst.is_contiguous() and src.is_contiguous()
    call("petr_upsample_add_tokens", ptr(dst), ptr(src), G, H, W, h, w, dst.shape[-1], stream_ptr(dst.device))
    return dst


def im2col3x3_tokens(x: torch.Tensor, G: int, H: int, W: int, mode: str = "f32") -> torch.Tensor:
    """fp32 tokens (G*H*W, C) -> (G*H*W, 9C) fp32 | bf16, or (G*H*W, 18C) bf16 [hi | lo] (mode 'split')."""
    _cuda(x, "x")
    assert x.is_contiguous() and x.dtype == torch.float32
    C = x.shape[-1]
    om = {"f32": 0, "bf16": 1, "split": 2}[mode]
    out = torch.empty((G * H * W, 9 * C * (2 if om == 2 else 1)),
                      dtype=torch.float32 if om == 0 else torch.bfloat16, device=x.device)
    call("petr_im2col3x3_tokens", ptr(x), ptr(out), om, G, H, W, C, stream_ptr(x.device))
    return out


def norm_scale(x: torch.Tensor, B: int, Q: int, along_queries: bool, power: float, eps: float, mult: float):
    _cuda(x, "x")
    assert x.is_contiguous() and x.dtype == torch.float32
    out = torch.empty_like(x)
    call("petr_norm_scale_f32", ptr(x), ptr(out), B, Q, x.shape[-1], int(bool(along_queries)), float(power),
         float(eps), float(mult), stream_ptr(x.device))
    return out


# ---- 3xTF32 tier: fp32-class GEMMs on tcgen05 (the fp32 parity path) --------------------------------------------
def split_rows_tf32(X: torch.Tensor, add=None) -> torch.Tensor:
    """fp32 (R,E) (+ add[r % add_rows]) -> fp32 (R,2E) [hi | lo], hi = tf32-rounded value, lo = remainder."""
    _cuda(X, "X")
    assert X.dim() == 2 and X.stride(1) == 1 and X.dtype == torch.float32
    R, E = X.shape
    out = torch.empty((R, 2 * E), dtype=torch.float32, device=X.device)
    if add is not None:
        assert add.is_contiguous() and add.shape[1] == E
    call("petr_split_rows_tf32", ptr(X), X.stride(0), ptr(add), add.shape[0] if add is not None else 0, ptr(out), R, E,
         stream_ptr(X.device))
    return out


def x3_weight_tf32(W: torch.Tensor) -> torch.Tensor:
    """fp32 (N,K) -> fp32 (N,3K) = [W_hi | W_hi | W_lo], the W operand of a 3xTF32 GEMM."""
    s = split_rows_tf32(W.contiguous())
    K = W.shape[1]
    return torch.cat([s[:, :K], s[:, :K], s[:, K:]], 1).contiguous()


def gemm_tf32x3(A_split: torch.Tensor, W3: torch.Tensor, bias=None, *, residual=None, act=ACT_NONE,
                split_out: bool = False, out=None, name: str = "gemm_tf32x3") -> torch.Tensor:
    """C = act(A @ W.T + bias) (+ residual) with A = [hi | lo] (M, 2K0) and W3 = x3_weight_tf32 (N, 3K0).
    split_out: fp32 (M, 2N) [hi | lo] rows instead of (M, N)."""
    _cuda(A_split, "A")
    assert A_split.dtype == torch.float32 and W3.dtype == torch.float32
    assert A_split.dim() == 2 and A_split.stride(1) == 1 and W3.is_contiguous()
    M = A_split.shape[0]
    N, K = W3.shape
    assert 3 * A_split.shape[1] == 2 * K, (A_split.shape, W3.shape)
    if out is None:
        out = torch.empty((M, 2 * N if split_out else N), dtype=torch.float32, device=A_split.device)
    ldc = residual.stride(0) if residual is not None else (N if split_out else out.stride(0))
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
        if not split_out:
            assert residual.stride(0) == out.stride(0)
    with _Timed(name):
        call("petr_gemm_tf32x3_tc", ptr(A_split), A_split.stride(0), A_split.shape[1], ptr(W3), W3.stride(0), ptr(bias),
             ptr(residual), act, ptr(out), ldc, 6 if split_out else 0, M, N, K, stream_ptr(A_split.device))
    return out
