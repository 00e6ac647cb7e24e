"""Training path of the head: the forward of petr_head.py:366-468 /
petrv2_head.py:429-540 in train mode with autograd, `loss` / `loss_single` / `get_targets` (petr_head.py:470-728) and
the Hungarian matching cost (core/bbox/assigners/hungarian_assigner_3d.py:117-143).

Every differentiable op is a `torch.autograd.Function` whose forward AND backward are kernels of libpetr_b200.so
(petr_attention_f32 / petr_layernorm_f32 forward, train_kernels.cu backward; the large Linear layers -- forward, input
and weight gradient -- as bf16x3 GEMMs on tcgen05, the ragged rest on the FFMA GEMMs).  PyTorch supplies
the tape, device memory, index bookkeeping of the matched targets and the optimizer; scipy's linear_sum_assignment does
the assignment on the host exactly as the reference does.  Dropout (attention probabilities, projection / FFN outputs)
uses a counter-based mask that the backward kernels re-evaluate, so nothing of size Lq x Lk is ever stored.
Activation checkpointing (`with_cp`, petr_transformer.py:201-212) only trades memory for recompute and is not needed at
180 GB; the attention itself already recomputes P in its backward.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import call, host_floats, ptr, stream_ptr
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID_GATE

_SEED = [0x5EED]


def manual_seed(seed: int):
    """Seeds the counter-based dropout masks of the training path."""
    _SEED[0] = int(seed)


def _next_seed() -> int:
    _SEED[0] = (_SEED[0] * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
    return _SEED[0]


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _transpose(W2: torch.Tensor) -> torch.Tensor:
    """(N,K) fp32 -> (K,N) through the layout kernel."""
    N, K = W2.shape
    return ops.tokens_to_nchw(_c(W2), 1, N).view(K, N)


# The large Linear layers run on the tensor cores as bf16x3 GEMMs (operands split [hi | lo], three bf16 products per fp32
# product, fp32 accumulation: ~1e-6 of the output scale, the tier the inference path uses for its query side); small or
# ragged shapes (the 256 -> 10 tails, NCHW input maps) stay on the FFMA kernel.
TC_GEMM = True
TC_MIN_ROWS = 256


def _tc_ok(M: int, N: int, K: int) -> bool:
    return TC_GEMM and M >= TC_MIN_ROWS and K % 64 == 0 and N % 8 == 0 and N >= 64


_SPLIT_CACHE: dict = {}          # shared operands only (linear(share_split=True)): key -> (A, A2, split); see _gemm_x3


def _gemm_x3(A, W2, bias, A2, residual, act, transpose_w=False, cache=False):
    """act((A + A2[m % rows]) W^T + bias) (+ residual) with W = W2 (or W2^T), fp32 in / out, on tcgen05: A is split
    to bf16 [hi | lo] by the library's split kernel, the weight packed [hi | hi | lo] by petr_x3_weight_bf16.
    cache: the split of a large forward activation is kept until the next forward starts -- the image tokens and
    their position embedding feed the K / V projections of all six decoder layers (the entry holds A, so its
    address cannot be reused while it is cached)."""
    A = _c(A)
    As = None
    key = ((A.data_ptr(), A._version, tuple(A.shape)) + ((A2.data_ptr(), A2._version) if A2 is not None else (0, 0))) if cache else None
    if key is not None:
        hit = _SPLIT_CACHE.get(key)
        if hit is not None:
            As = hit[2]
    if As is None:
        As = ops.split_rows(A, add=A2)
        if key is not None:
            _SPLIT_CACHE[key] = (A, A2, As)
    W3 = ops.x3_weight_kernel(W2, transpose_w)
    return ops.gemm_bf16_tc(As, W3, bias, residual=residual, act=act, x3=True)


def _tc_dw_ok(M: int, N: int, K: int) -> bool:
    return TC_GEMM and M >= TC_MIN_ROWS and N >= 64 and K >= 64 and N % 8 == 0 and K % 8 == 0


def _dw_x3(dZ, A, A2, share=False):
    """dW (N,K) = dZ^T (A + A2[m % rows]) on tcgen05: both factors transposed and split by the library
    (petr_split_transpose_bf16, rows padded to a multiple of 64 with zeros; the transposed image tokens / position
    embedding are shared by the K / V projections of all six layers through the same cache as their forward split), one
    bf16x3 GEMM over the row axis cut into split-K parts so that ~148 CTAs run, parts summed by petr_sum_parts_f32."""
    M, N = dZ.shape
    K = A.shape[1]
    Mp = (M + 63) // 64 * 64
    Yt = ops.split_transpose(dZ, None, Mp, False)                   # (N, 2 Mp)
    At = None
    key = ("T", A.data_ptr(), A._version, tuple(A.shape)) + ((A2.data_ptr(), A2._version) if A2 is not None else (0, 0))
    if share:
        hit = _SPLIT_CACHE.get(key)
        if hit is not None:
            At = hit[2]
    if At is None:
        At = ops.split_transpose(A, A2, Mp, True)                   # (K, 3 Mp)
        if share:
            _SPLIT_CACHE[key] = (A, A2, At)
    num_k = 3 * Mp // 64
    tiles = ((N + 127) // 128) * ((K + 255) // 256)
    ks = max(1, min(148 // tiles, num_k // 4))
    parts = ops.gemm_bf16_tc(Yt, At, None, x3=True, ksplit=ks, name="gemm_bf16_tc_dw")
    if ks == 1:
        return parts
    dW = torch.empty((N, K), dtype=torch.float32, device=dZ.device)
    call("petr_sum_parts_f32", ptr(parts), ks, N * K, ptr(dW), stream_ptr(dZ.device))
    return dW


# ------------------------------------------------------------------------------------------------------------
# autograd wrappers
# ------------------------------------------------------------------------------------------------------------
class _Linear(Function):
    """y = act((A + A2[m % rows]) W^T + b) (+ residual); A row-major (M,K) or an NCHW map read as tokens."""

    @staticmethod
    def forward(ctx, A, W, bias, A2, residual, act, nchw_hw, share_split=False):
        W2 = W.detach().reshape(W.shape[0], -1)
        assert not (act != ACT_NONE and residual is not None)
        A = _c(A.detach())
        if _tc_ok(A.shape[0], W2.shape[0], W2.shape[1]) and nchw_hw == 0:
            out = _gemm_x3(A, _c(W2), bias.detach() if bias is not None else None,
                           A2.detach() if A2 is not None else None,
                           _c(residual.detach()) if residual is not None else None, act, cache=share_split)
        else:
            out = ops.gemm_f32(A, _c(W2), bias.detach() if bias is not None else None,
                               A2=A2.detach() if A2 is not None else None,
                               residual=residual.detach() if residual is not None else None, act=act, nchw_hw=nchw_hw)
        ctx.act, ctx.nchw_hw, ctx.wshape, ctx.share_split = act, nchw_hw, W.shape, share_split
        ctx.save_for_backward(A, W2, A2.detach() if A2 is not None else None, out if act == ACT_RELU else None)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dOut):
        A, W2, A2, out = ctx.saved_tensors
        dOut = _c(dOut)
        dev = dOut.device
        M, N = dOut.shape
        K = W2.shape[1]
        dZ = dOut
        if ctx.act == ACT_RELU:
            dZ = torch.empty_like(dOut)
            call("petr_relu_bwd_f32", ptr(dOut), ptr(out), ptr(dZ), dOut.numel(), stream_ptr(dev))
        need = ctx.needs_input_grad
        dA = dW = db = dA2 = dres = None
        if need[0] or (A2 is not None and need[3]):
            if _tc_ok(M, K, N):
                dA_tok = _gemm_x3(dZ, W2, None, None, None, ACT_NONE, transpose_w=True)
            else:
                dA_tok = ops.gemm_f32(dZ, _transpose(W2))                   # (M, K) = dZ @ W
            if need[0]:
                if ctx.nchw_hw > 0:
                    G = M // ctx.nchw_hw
                    dA = ops.tokens_to_nchw(dA_tok, G, ctx.nchw_hw).view(A.shape)
                else:
                    dA = dA_tok
            if A2 is not None and need[3]:
                dA2 = torch.zeros_like(A2)
                call("petr_rowsum_periodic_f32", ptr(dA_tok), M, K, A2.shape[0], ptr(dA2), stream_ptr(dev))
        if need[1] and ctx.nchw_hw == 0 and _tc_dw_ok(M, N, K):
            dW = _dw_x3(dZ, A, A2, ctx.share_split).view(ctx.wshape)
        elif need[1]:
            dW = torch.zeros((N, K), dtype=torch.float32, device=dev)
            if ctx.nchw_hw > 0:
                call("petr_gemm_f32_tn", ptr(dZ), dZ.stride(0), ptr(A), K, ops.A_NCHW, ctx.nchw_hw, None, 0, ptr(dW),
                     M, N, K, stream_ptr(dev))
            else:
                call("petr_gemm_f32_tn", ptr(dZ), dZ.stride(0), ptr(A), A.stride(0), ops.A_ROWMAJOR, 0, ptr(A2),
                     A2.shape[0] if A2 is not None else 0, ptr(dW), M, N, K, stream_ptr(dev))
            dW = dW.view(ctx.wshape)
        if ctx.has_bias and need[2]:
            db = torch.zeros((N,), dtype=torch.float32, device=dev)
            call("petr_colsum_f32", ptr(dZ), dZ.stride(0), M, N, ptr(db), stream_ptr(dev))
        if need[4]:
            dres = dOut
        return dA, dW, db, dA2, dres, None, None, None


def linear(A, W, bias=None, A2=None, residual=None, act=ACT_NONE, nchw_hw=0, share_split=False):
    """share_split: A (+ A2) feeds several Linear layers of one step (the image tokens and their position embedding
    under the K / V projections of all decoder layers): keep its bf16 split / transposed split until the next forward."""
    return _Linear.apply(A, W, bias, A2, residual, act, nchw_hw, share_split)


class _LinearGate(Function):
    """out = gate * sigmoid(A W^T + b) (+ residual): PETRv2 SELayer tail (petrv2_head.py:48-60)."""

    @staticmethod
    def forward(ctx, A, W, bias, gate, residual):
        W2 = _c(W.detach().reshape(W.shape[0], -1))
        A, gate = _c(A.detach()), _c(gate.detach())
        z = ops.gemm_f32(A, W2, bias.detach())
        out = ops.gemm_f32(A, W2, bias.detach(), act=ACT_SIGMOID_GATE, gate=gate,
                           residual=residual.detach() if residual is not None else None)
        ctx.wshape = W.shape
        ctx.save_for_backward(A, W2, z, gate)
        return out

    @staticmethod
    def backward(ctx, dOut):
        A, W2, z, gate = ctx.saved_tensors
        dOut = _c(dOut)
        dev = dOut.device
        M, N = dOut.shape
        K = W2.shape[1]
        dZ, dGate = torch.empty_like(dOut), torch.empty_like(dOut)
        call("petr_sigmoid_gate_bwd_f32", ptr(dOut), ptr(z), ptr(gate), ptr(dZ), ptr(dGate), dOut.numel(), stream_ptr(dev))
        dA = ops.gemm_f32(dZ, _transpose(W2))
        dW = torch.zeros((N, K), dtype=torch.float32, device=dev)
        call("petr_gemm_f32_tn", ptr(dZ), N, ptr(A), A.stride(0), ops.A_ROWMAJOR, 0, None, 0, ptr(dW), M, N, K, stream_ptr(dev))
        db = torch.zeros((N,), dtype=torch.float32, device=dev)
        call("petr_colsum_f32", ptr(dZ), N, M, N, ptr(db), stream_ptr(dev))
        return dA, dW.view(ctx.wshape), db, dGate, dOut if ctx.needs_input_grad[4] else None


class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, X, gamma, beta, eps, relu, nan_to_num):
        X = _c(X.detach())
        out = ops.layernorm(X, gamma.detach(), beta.detach(), eps, relu=relu, nan_to_num=nan_to_num)
        ctx.cfg = (eps, relu, nan_to_num)
        ctx.save_for_backward(X, gamma.detach(), beta.detach())
        return out

    @staticmethod
    def backward(ctx, dY):
        X, gamma, beta = ctx.saved_tensors
        eps, relu, n2n = ctx.cfg
        dY = _c(dY)
        dev = dY.device
        E = X.shape[-1]
        rows = X.numel() // E
        dX = torch.empty_like(X)
        dg, db = torch.zeros_like(gamma), torch.zeros_like(beta)
        call("petr_layernorm_bwd_f32", ptr(X), ptr(gamma), ptr(beta), ptr(dY), ptr(dX), ptr(dg), ptr(db), rows, E,
             float(eps), int(relu), int(n2n), stream_ptr(dev))
        return dX, dg, db, None, None, None


def layernorm(X, gamma, beta, eps=1e-5, relu=False, nan_to_num=False):
    return _LayerNorm.apply(X, gamma, beta, eps, relu, nan_to_num)


class _Attention(Function):
    """softmax(Q K^T / sqrt(d) + mask) V with attention-probability dropout; Q/K/V may be column slices."""

    @staticmethod
    def forward(ctx, Q, K, V, key_mask, B, Lq, Lk, heads, drop_p, seed):
        Q, K, V = Q.detach(), K.detach(), V.detach()
        lse = torch.empty((B, heads, Lq), dtype=torch.float32, device=Q.device)
        O = ops.attention_f32(Q, K, V, key_mask, B, Lq, Lk, heads, lse=lse, drop_p=drop_p, seed=seed)
        ctx.cfg = (B, Lq, Lk, heads, drop_p, seed)
        ctx.key_mask = key_mask
        ctx.save_for_backward(Q, K, V, O, lse)
        return O

    @staticmethod
    def backward(ctx, dO):
        Q, K, V, O, lse = ctx.saved_tensors
        B, Lq, Lk, heads, drop_p, seed = ctx.cfg
        dO = _c(dO)
        dev = dO.device
        E = heads * 32
        dQ = torch.empty((B * Lq, E), dtype=torch.float32, device=dev)
        dK = torch.empty((B * Lk, E), dtype=torch.float32, device=dev)
        dV = torch.empty((B * Lk, E), dtype=torch.float32, device=dev)
        call("petr_attention_f32_bwd", ptr(Q), Q.stride(0), ptr(K), K.stride(0), ptr(V), V.stride(0), ptr(O), O.stride(0),
             ptr(dO), dO.stride(0), ptr(lse), ptr(ctx.key_mask), None, 0, ptr(dQ), E, ptr(dK), E, ptr(dV), E,
             float(drop_p), int(seed), B, Lq, Lk, heads, 32, 1.0 / math.sqrt(32.0), stream_ptr(dev))
        return dQ, dK, dV, None, None, None, None, None, None, None


class _Dropout(Function):
    """y = dropout(x) (+ residual)."""

    @staticmethod
    def forward(ctx, X, residual, p, seed):
        X = _c(X.detach())
        Y = torch.empty_like(X)
        res = _c(residual.detach()) if residual is not None else None      # (held: a temporary must outlive the launch setup)
        call("petr_dropout_f32", ptr(X), ptr(res), ptr(Y), X.numel(), float(p), int(seed), stream_ptr(X.device))
        ctx.cfg = (p, seed)
        return Y

    @staticmethod
    def backward(ctx, dY):
        p, seed = ctx.cfg
        dY = _c(dY)
        dX = torch.empty_like(dY)
        call("petr_dropout_f32", ptr(dY), None, ptr(dX), dY.numel(), float(p), int(seed), stream_ptr(dY.device))
        return dX, (dY if ctx.needs_input_grad[1] else None), None, None


def dropout(X, p: float, residual=None):
    """nn.Dropout(p) in training (+ fused residual); p == 0 with a residual is handled by the caller's fused GEMM."""
    if p <= 0.0:
        assert residual is None
        return X
    return _Dropout.apply(X, residual, p, _next_seed())


class _NormScale(Function):
    """mmdet NormedLinear operand scaling (petr_head.py:231-232): x * mult / (||x||^power + eps), norm over the query
    axis of (B, Q, E) rows (along_queries) or over each row (the weight)."""

    @staticmethod
    def forward(ctx, x, B, Q, along_queries, power, eps, mult):
        x = _c(x.detach())
        ctx.cfg = (B, Q, bool(along_queries), float(power), float(eps), float(mult))
        ctx.save_for_backward(x)
        return ops.norm_scale(x, B, Q, along_queries, power, eps, mult)

    @staticmethod
    def backward(ctx, dOut):
        (x,) = ctx.saved_tensors
        B, Q, along, power, eps, mult = ctx.cfg
        dOut = _c(dOut)
        dx = torch.empty_like(x)
        call("petr_norm_scale_bwd_f32", ptr(x), ptr(dOut), ptr(dx), B, Q, x.shape[-1], int(along), power, eps, mult,
             stream_ptr(x.device))
        return dx, None, None, None, None, None, None


class _LearnedPos(Function):
    """LearnedPositionalEncoding3D.forward as tokens (positional_encoding.py:137-160): (G*H*W, 2F) =
    [col_embed[w] | row_embed[h]]; the backward sums the token gradients into the two tables."""

    @staticmethod
    def forward(ctx, row_w, col_w, G, H, W):
        row_w, col_w = _c(row_w.detach().float()), _c(col_w.detach().float())
        ctx.cfg = (G, H, W, tuple(row_w.shape), tuple(col_w.shape))
        return ops.learned_pos_features(row_w, col_w, G, H, W)

    @staticmethod
    def backward(ctx, dOut):
        G, H, W, rs, cs = ctx.cfg
        dOut = _c(dOut)
        d_row = torch.zeros(rs, dtype=torch.float32, device=dOut.device)
        d_col = torch.zeros(cs, dtype=torch.float32, device=dOut.device)
        call("petr_learned_pos_bwd", ptr(dOut), G, H, W, rs[1], ptr(d_row), ptr(d_col), stream_ptr(dOut.device))
        return d_row, d_col, None, None, None


class _Pos2PosEmb(Function):
    @staticmethod
    def forward(ctx, ref, dim_t):
        ref = _c(ref.detach())
        ctx.save_for_backward(ref, dim_t)
        return ops.pos2posemb3d(ref, dim_t)

    @staticmethod
    def backward(ctx, dOut):
        ref, dim_t = ctx.saved_tensors
        dref = torch.zeros_like(ref)
        dOut = _c(dOut)
        call("petr_pos2posemb3d_bwd", ptr(ref), ptr(dim_t), ptr(dOut), ref.shape[0], dim_t.numel(), ptr(dref),
             stream_ptr(ref.device))
        return dref, None


class _DecodeBoxes(Function):
    """petr_head.py:440-460: reference add, sigmoid, pc_range affine (and PETRv2's velocity / mean time stamp)."""

    @staticmethod
    def forward(ctx, reg, ref, pc_range, time_div, L, B, Q):
        out = reg.detach().clone()
        ref = _c(ref.detach())
        ops.decode_boxes(out, ref, pc_range, time_div, L, B, Q)
        ctx.cfg = (list(pc_range), L, B, Q)
        ctx.time_div = time_div
        ctx.save_for_backward(out, ref)
        return out

    @staticmethod
    def backward(ctx, dOut):
        out, ref = ctx.saved_tensors
        pc, L, B, Q = ctx.cfg
        dOut = _c(dOut)
        dReg = torch.empty_like(dOut)
        dref = torch.zeros_like(ref)
        call("petr_decode_boxes_bwd", ptr(out), ptr(dOut), ptr(ref), host_floats(pc), ptr(ctx.time_div), ptr(dReg), ptr(dref),
             L, B, Q, out.shape[-1], stream_ptr(out.device))
        return dReg, dref, None, None, None, None, None


class _FocalLoss(Function):
    """loss = scale * sum_i w_i FL(x_i, label_i)  (mmdet FocalLoss with avg_factor and loss_weight folded into scale)."""

    @staticmethod
    def forward(ctx, X, labels, weights, alpha, gamma, scale):
        X = _c(X.detach())
        rows, C = X.shape
        acc = torch.zeros((1,), dtype=torch.float32, device=X.device)
        dX = torch.empty_like(X)
        call("petr_focal_loss", ptr(X), ptr(labels), ptr(weights), rows, C, float(alpha), float(gamma), float(scale),
             ptr(acc), ptr(dX), stream_ptr(X.device))
        ctx.save_for_backward(dX)
        return acc[0] * scale

    @staticmethod
    def backward(ctx, g):
        (dX,) = ctx.saved_tensors
        return dX * g, None, None, None, None, None


class _L1Loss(Function):
    @staticmethod
    def forward(ctx, X, T, W, scale):
        X = _c(X.detach())
        rows, code = X.shape
        acc = torch.zeros((1,), dtype=torch.float32, device=X.device)
        dX = torch.empty_like(X)
        T, W = _c(T), _c(W)
        call("petr_l1_loss", ptr(X), ptr(T), ptr(W), rows, code, float(scale), ptr(acc), ptr(dX), stream_ptr(X.device))
        ctx.save_for_backward(dX)
        return acc[0] * scale

    @staticmethod
    def backward(ctx, g):
        (dX,) = ctx.saved_tensors
        return dX * g, None, None, None


# ------------------------------------------------------------------------------------------------------------
# forward in train mode
# ------------------------------------------------------------------------------------------------------------
class TrainEngine:
    """Differentiable forward of the head on the fp32 kernels.  Parameters are read from the module (fp32, CUDA)."""

    def __init__(self, head):
        from .engine import HeadEngine, sine_dim_t
        self.head = head
        self.eng = HeadEngine(head)                 # spec + the gradient-free kernels (masks, frustum, encodings)
        self._consts = {}

    def _constants(self, dev):
        from .engine import sine_dim_t
        key = str(dev)
        if key not in self._consts:
            s = self.eng.spec
            c = dict(dim_t_query=sine_dim_t(128, 10000).to(dev))
            if s.sine is not None:
                c["dim_t_sine"] = sine_dim_t(s.sine["num_feats"], s.sine["temperature"]).to(dev)
            self._consts[key] = c
        return self._consts[key]

    # ---- bricks ----
    def _mha(self, att, X, qpos, Kin, Kpos, Vin, kmask, B, Lq, Lk, training):
        """PETRMultiheadAttention / mmcv MultiheadAttention (petr_transformer.py:280-367) on (B*L, E) rows:
        identity + dropout_layer(proj_drop(out_proj(attn(q + qpos, k + kpos, v))))."""
        E = att.embed_dims
        w, b = att.attn.in_proj_weight, att.attn.in_proj_bias
        q = linear(X, w[:E], b[:E], A2=qpos)
        shared = Kin is not X                           # cross-attention: the same tokens under every layer's K / V
        k = linear(Kin, w[E:2 * E], b[E:2 * E], A2=Kpos, share_split=shared)
        v = linear(Vin, w[2 * E:], b[2 * E:], share_split=shared)
        p_attn = float(att.attn.dropout) if training else 0.0
        o = _Attention.apply(q, k, v, kmask, B, Lq, Lk, att.num_heads, p_attn, _next_seed() if p_attn > 0 else 0)
        p_proj = float(att.proj_drop.p) if training else 0.0
        p_layer = float(att.dropout_layer.p) if (training and isinstance(att.dropout_layer, torch.nn.Dropout)) else 0.0
        ow, ob = att.attn.out_proj.weight, att.attn.out_proj.bias
        if p_proj == 0.0 and p_layer == 0.0:
            return linear(o, ow, ob, residual=X)
        y = dropout(linear(o, ow, ob), p_proj)
        if p_layer > 0.0:
            return dropout(y, p_layer, residual=X)
        return _Dropout.apply(y, X, 0.0, 0)            # p == 0: plain residual add through the same kernel

    def _ffn(self, ffn, X, training):
        """mmcv FFN: Linear, ReLU, Dropout, Linear, Dropout, + identity (dropout_layer: Identity in the PETR configs)."""
        assert ffn.num_fcs == 2 and ffn.add_identity
        l1, d1 = ffn.layers[0][0], ffn.layers[0][2]
        l2, d2 = ffn.layers[1], ffn.layers[2]
        p1 = float(d1.p) if training else 0.0
        p2 = float(d2.p) if training else 0.0
        p3 = float(ffn.dropout_layer.p) if (training and isinstance(ffn.dropout_layer, torch.nn.Dropout)) else 0.0
        h = dropout(linear(X, l1.weight, l1.bias, act=ACT_RELU), p1)
        if p2 == 0.0 and p3 == 0.0:
            return linear(h, l2.weight, l2.bias, residual=X)
        y = linear(h, l2.weight, l2.bias)
        if p3 > 0.0:
            return dropout(dropout(y, p2), p3, residual=X)
        return dropout(y, p2, residual=X)

    def _ln(self, ln, X, relu=False, nan_to_num=False):
        return layernorm(X, ln.weight, ln.bias, ln.eps, relu, nan_to_num)

    def _mlp1x1(self, seq, x, residual=None, nchw_hw=0):
        h = linear(x, seq[0].weight, seq[0].bias, act=ACT_RELU, nchw_hw=nchw_hw)
        return linear(h, seq[2].weight, seq[2].bias, residual=residual)

    # ---- the head ----
    def forward(self, x: torch.Tensor, img_metas, training: bool = True) -> dict:
        from .engine import host_prep, stage_meta
        head = self.head
        eng = self.eng
        s = eng.spec
        if not x.is_cuda:
            raise RuntimeError("petr_b200: features must be CUDA tensors (no CPU fallback)")
        _SPLIT_CACHE.clear()
        x = x.float()
        B, N, Cin, H, W = x.shape
        dev = x.device
        Lk, E, Lq = N * H * W, s.E, s.Lq
        mi = stage_meta(host_prep(img_metas, N, H, W, s.with_time), dev)
        P = self._constants(dev)                            # dim_t tables; parameters are read from the module
        masks = eng.masks(mi, H, W)
        kmask = masks.view(B, Lk)
        # key side (petr_head.py:381-420)
        mem = linear(x.contiguous().view(B * N, Cin, H, W), head.input_proj.weight, head.input_proj.bias, nchw_hw=H * W)
        if s.pos_kind == "learned":
            pe = head.positional_encoding
            sf = _LearnedPos.apply(pe.row_embed.weight, pe.col_embed.weight, B * N, H, W)
        else:
            sf = eng.pos_features(P, masks)                 # no gradient: depends on the masks only
        if not s.with_position and not s.with_multiview:
            pos = sf
        else:
            pos = None
            if s.with_position:
                feat, _, _ = eng.frustum(mi, B, N, H, W, masks, want_mask=False)          # no gradient (img_metas only)
                pos = self._mlp1x1(head.position_encoder, feat)
                if s.with_fpe:
                    se = linear(mem, head.fpe.conv_reduce.weight, head.fpe.conv_reduce.bias, act=ACT_RELU)
                    pos = _LinearGate.apply(se, head.fpe.conv_expand.weight, head.fpe.conv_expand.bias, pos, None)
            pos = self._mlp1x1(head.adapt_pos3d, sf, residual=pos)
        # query side
        ref = head.reference_points.weight
        qe = _Pos2PosEmb.apply(ref, P["dim_t_query"])
        qpos = linear(linear(qe, head.query_embedding[0].weight, head.query_embedding[0].bias, act=ACT_RELU),
                      head.query_embedding[2].weight, head.query_embedding[2].bias)               # (Lq, E)
        dec = head.transformer.decoder
        X = torch.zeros((B * Lq, E), dtype=torch.float32, device=dev)
        outs = []
        for layer in dec.layers:
            sa, ca = layer.attentions
            X = self._ln(layer.norms[0], self._mha(sa, X, qpos, X, qpos, X, None, B, Lq, Lq, training))
            X = self._ln(layer.norms[1], self._mha(ca, X, qpos, mem, pos, mem, kmask, B, Lq, Lk, training))
            X = self._ln(layer.norms[2], self._ffn(layer.ffns[0], X, training))
            outs.append(self._ln(dec.post_norm, X, nan_to_num=True))          # post_norm on a copy + nan_to_num
        L = len(outs)
        cls_all, reg_all = [], []
        for l in range(L):
            cb, rb = head.cls_branches[l], head.reg_branches[l]
            h = outs[l]
            for i in range(s.num_reg_fcs):
                h = self._ln(cb[3 * i + 1], linear(h, cb[3 * i].weight, cb[3 * i].bias), relu=True)
            last = cb[3 * s.num_reg_fcs]
            if s.normedlinear:                              # mmdet NormedLinear (petr_head.py:231-234)
                wn = _NormScale.apply(last.weight, 1, last.weight.shape[0], False, last.power, last.eps, 1.0)
                h = _NormScale.apply(h, B, Lq, True, last.power, last.eps, last.tempearture)
                cls_all.append(linear(h, wn, last.bias))
            else:
                cls_all.append(linear(h, last.weight, last.bias))
            h = outs[l]
            if s.with_multi:
                for i in range(s.num_reg_fcs):
                    h = linear(h, rb.reg_branch[3 * i].weight, rb.reg_branch[3 * i].bias, act=ACT_RELU)
                parts = []
                for th in rb.task_heads:
                    parts.append(linear(linear(h, th[0].weight, th[0].bias, act=ACT_RELU), th[2].weight, th[2].bias))
                reg_all.append(torch.cat(parts, 1))
            else:
                for i in range(s.num_reg_fcs):
                    h = linear(h, rb[2 * i].weight, rb[2 * i].bias, act=ACT_RELU)
                reg_all.append(linear(h, rb[2 * s.num_reg_fcs].weight, rb[2 * s.num_reg_fcs].bias))
        cls = torch.stack(cls_all).view(L, B, Lq, s.cls_out)
        reg = torch.stack(reg_all).view(L * B * Lq, s.code_size)
        box = _DecodeBoxes.apply(reg, ref, s.pc_range, mi.time_div, L, B, Lq).view(L, B, Lq, s.code_size)
        return dict(all_cls_scores=cls, all_bbox_preds=box, enc_cls_scores=None, enc_bbox_preds=None)


# ------------------------------------------------------------------------------------------------------------
# loss (petr_head.py:470-728)
# ------------------------------------------------------------------------------------------------------------
def _gt_tensor(gt, device) -> torch.Tensor:
    """LiDARInstance3DBoxes-like (gravity_center + tensor[:, 3:], petr_head.py:690-692) or a plain
    (G, 9) tensor already in (cx, cy, cz, w, l, h, rot, vx, vy) gravity-centre form."""
    if hasattr(gt, "gravity_center"):
        return torch.cat((gt.gravity_center, gt.tensor[:, 3:]), dim=1).to(device=device, dtype=torch.float32)
    return gt.to(device=device, dtype=torch.float32)


def _reduce_mean(t: torch.Tensor) -> torch.Tensor:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return t
    t = t.clone()
    dist.all_reduce(t.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
    return t


class HeadLoss:
    """`loss` of PETRHead / PETRv2Head with HungarianAssigner3D (FocalLossCost + BBox3DL1Cost; the IoU cost has weight
    0 in every config), FocalLoss and L1Loss as configured by `loss_cls` / `loss_bbox` / `train_cfg.assigner`."""

    def __init__(self, head):
        self.head = head
        lc = head.loss_cls.cfg or {}
        lb = head.loss_bbox.cfg or {}
        if lc.get("type") != "FocalLoss" or not lc.get("use_sigmoid", False):
            raise NotImplementedError("petr_b200 loss: loss_cls must be a sigmoid FocalLoss (every PETR config)")
        if lb.get("type") != "L1Loss":
            raise NotImplementedError("petr_b200 loss: loss_bbox must be L1Loss (every PETR config)")
        self.alpha, self.gamma = float(lc.get("alpha", 0.25)), float(lc.get("gamma", 2.0))
        self.w_cls, self.w_bbox = float(lc.get("loss_weight", 1.0)), float(lb.get("loss_weight", 1.0))
        tc = head.train_cfg or {}
        tc = tc.get("pts", tc) if isinstance(tc, dict) else {}
        ac = (tc or {}).get("assigner", None) or dict(cls_cost=dict(type='FocalLossCost', weight=2.0),
                                                       reg_cost=dict(type='BBox3DL1Cost', weight=0.25))
        cc, rc = ac.get("cls_cost", {}), ac.get("reg_cost", {})
        if cc.get("type", "FocalLossCost") != "FocalLossCost" or rc.get("type", "BBox3DL1Cost") != "BBox3DL1Cost":
            raise NotImplementedError("petr_b200 loss: assigner costs must be FocalLossCost + BBox3DL1Cost")
        self.cost_cls_w, self.cost_reg_w = float(cc.get("weight", 1.0)), float(rc.get("weight", 1.0))
        self.cost_alpha, self.cost_gamma, self.cost_eps = float(cc.get("alpha", 0.25)), float(cc.get("gamma", 2.0)), \
            float(cc.get("eps", 1e-12))

    # ---- assignment (hungarian_assigner_3d.py:60-143), all decoder layers and images with one host sync ----
    def assign(self, cls_all, box_all, gts, labels):
        from scipy.optimize import linear_sum_assignment
        L, B, Q, C = cls_all.shape
        code = box_all.shape[-1]
        dev = cls_all.device
        cls_d, box_d = _c(cls_all.detach()), _c(box_all.detach())
        gmax = max([int(g.shape[0]) for g in gts] + [1])
        cost = torch.zeros((L, B, Q, gmax), dtype=torch.float32, device=dev)
        ngts = []
        for b in range(B):
            G = int(gts[b].shape[0])
            if G == 0:
                ngts.append(None)
                continue
            ng = torch.empty((G, gts[b].shape[1] + 1), dtype=torch.float32, device=dev)
            gb = _c(gts[b])
            call("petr_normalize_bbox", ptr(gb), G, gts[b].shape[1], ptr(ng), stream_ptr(dev))
            ngts.append(ng)
        tmp = {}
        for l in range(L):
            for b in range(B):
                if ngts[b] is None:
                    continue
                G = ngts[b].shape[0]
                c = torch.empty((Q, G), dtype=torch.float32, device=dev)
                lb_ = _c(labels[b].long())
                call("petr_hungarian_cost", ptr(cls_d[l, b]), ptr(box_d[l, b]), ptr(lb_), ptr(ngts[b]), Q, G, C,
                     code, ngts[b].shape[1], self.cost_cls_w, self.cost_reg_w, self.cost_alpha, self.cost_gamma,
                     self.cost_eps, ptr(c), stream_ptr(dev))
                cost[l, b, :, :G] = c
        cost_h = cost.cpu().numpy()                               # the one synchronisation of the loss
        matches = []
        for l in range(L):
            row = []
            for b in range(B):
                G = 0 if ngts[b] is None else int(ngts[b].shape[0])
                if G == 0:
                    row.append((np.zeros((0,), np.int64), np.zeros((0,), np.int64)))
                else:
                    r, cidx = linear_sum_assignment(cost_h[l, b, :, :G])
                    row.append((r.astype(np.int64), cidx.astype(np.int64)))
            matches.append(row)
        return matches

    def loss_single(self, cls_scores, bbox_preds, gts, labels, matches):
        """petr_head.py:578-645 for one decoder layer."""
        head = self.head
        B, Q, C = cls_scores.shape
        code = bbox_preds.shape[-1]
        dev = cls_scores.device
        lab = torch.full((B * Q,), head.num_classes, dtype=torch.int64, device=dev)
        gcode = gts[0].shape[1] if len(gts) else 9
        tgt = torch.zeros((B * Q, gcode), dtype=torch.float32, device=dev)
        wts = torch.zeros((B * Q, code), dtype=torch.float32, device=dev)
        cw = head.code_weights.detach().to(dev)
        num_pos = 0
        for b in range(B):
            r, cidx = matches[b]
            if len(r) == 0:
                continue
            rr = torch.from_numpy(r).to(dev) + b * Q
            cc = torch.from_numpy(cidx).to(dev)
            lab[rr] = labels[b].long()[cc]
            tgt[rr] = gts[b][cc]
            wts[rr] = cw
            num_pos += len(r)
        num_neg = B * Q - num_pos
        cls_avg = num_pos * 1.0 + num_neg * head.bg_cls_weight
        if head.sync_cls_avg_factor:
            cls_avg = float(_reduce_mean(torch.tensor([cls_avg], dtype=torch.float32, device=dev)).item())
        cls_avg = max(cls_avg, 1)
        loss_cls = _FocalLoss.apply(cls_scores.reshape(B * Q, C), lab, None, self.alpha, self.gamma, self.w_cls / cls_avg)
        npos = float(torch.clamp(_reduce_mean(torch.tensor([float(num_pos)], device=dev)), min=1).item())
        ntgt = torch.empty((B * Q, gcode + 1), dtype=torch.float32, device=dev)
        call("petr_normalize_bbox", ptr(tgt), B * Q, gcode, ptr(ntgt), stream_ptr(dev))     # unmatched rows: log 0 = -inf
        loss_bbox = _L1Loss.apply(bbox_preds.reshape(B * Q, code)[:, :10], ntgt[:, :10], wts[:, :10], self.w_bbox / npos)
        return torch.nan_to_num(loss_cls), torch.nan_to_num(loss_bbox)

    def __call__(self, gt_bboxes_list, gt_labels_list, preds_dicts, gt_bboxes_ignore=None):
        assert gt_bboxes_ignore is None, f'{type(self.head).__name__} only supports for gt_bboxes_ignore setting to None.'
        cls_all, box_all = preds_dicts['all_cls_scores'], preds_dicts['all_bbox_preds']
        if preds_dicts.get('enc_cls_scores') is not None:
            raise NotImplementedError("two-stage (enc_*) losses are never produced by the PETR heads")
        dev = cls_all.device
        gts = [_gt_tensor(g, dev) for g in gt_bboxes_list]
        labels = [t.to(dev) for t in gt_labels_list]
        matches = self.assign(cls_all, box_all, gts, labels)
        L = cls_all.shape[0]
        out = {}
        for l in range(L):
            lc, lb = self.loss_single(cls_all[l], box_all[l], gts, labels, matches[l])
            pre = "" if l == L - 1 else f"d{l}."
            out[pre + "loss_cls"], out[pre + "loss_bbox"] = lc, lb
        # the reference orders the dict: last layer first, then d0 ... d{L-2}
        return {k: out[k] for k in (["loss_cls", "loss_bbox"] +
                                    [f"d{l}.{n}" for l in range(L - 1) for n in ("loss_cls", "loss_bbox")])}
