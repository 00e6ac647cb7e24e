"""bf16 / tcgen05 path of the head engine (see engine.py): the Lk-side work -- position_encoder
and adapt_pos3d MLPs, the K/V projection of all six decoder layers and the flash cross-attention
-- runs on the tensor cores with bf16 operands and fp32 accumulation.  Query-side GEMMs, the
residual stream, LayerNorm, softmax statistics and the branches stay fp32 (SURVEY.md hard part 3:
cx/cy in metres are the tight output channels under the 2e-2 tolerance)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import ops
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID_GATE

NUM_SMS = 148


def _nearest_src(dst: int, in_size: int, out_size: int) -> int:
    # same rule as the device kernel / torch upsample_nearest (fp32 arithmetic)
    if out_size == in_size:
        return dst
    if out_size == 2 * in_size:
        return dst >> 1
    scale = np.float32(in_size) / np.float32(out_size)
    return min(int(np.floor(np.float32(dst) * scale)), in_size - 1)


def any_key_masked(img_metas, N: int, H: int, W: int) -> bool:
    """Host-side exact test whether petr_head.py:383-394 masks any key (no device sync)."""
    pad_h, pad_w = img_metas[0]["pad_shape"][0][:2]
    last_r, last_c = _nearest_src(H - 1, pad_h, H), _nearest_src(W - 1, pad_w, W)
    for m in img_metas:
        for n in range(N):
            ih, iw = m["img_shape"][n][:2]
            if not (last_r < ih and last_c < iw):
                return True
    return False


def choose_nsplit(B: int, heads: int, Lq: int, Lk: int) -> int:
    """Key splits so that the grid fills the 148 SMs in whole waves (1 CTA per SM)."""
    base = B * heads * ((Lq + 127) // 128)
    tiles = (Lk + 127) // 128
    best, best_eff = 1, 0.0
    for ns in range(1, min(16, tiles) + 1):
        ctas = base * ns
        eff = ctas / (((ctas + NUM_SMS - 1) // NUM_SMS) * NUM_SMS)
        if tiles // ns < 4:
            break
        if eff > best_eff + 0.03:
            best, best_eff = ns, eff
    return best


def packed_tc(eng, device) -> dict:
    P = eng.packed(device)
    if "tc" in P:
        return P["tc"]
    s = eng.spec
    E = s.E
    bf = torch.bfloat16
    T = dict()
    if s.with_position:
        T["pe_w1"] = P["pe_w1"].to(bf).contiguous()
        T["pe_w2"] = P["pe_w2"].to(bf).contiguous()
    T["ad_w1"] = P["ad_w1"].to(bf).contiguous()
    T["ad_w2"] = P["ad_w2"].to(bf).contiguous()
    T["wk"] = torch.cat([lw["ca_in_w"][E:2 * E] for lw in P["layers"]], 0).to(bf).contiguous()
    T["bk"] = torch.cat([lw["ca_in_b"][E:2 * E] for lw in P["layers"]], 0).contiguous()
    T["wv"] = torch.cat([lw["ca_in_w"][2 * E:] for lw in P["layers"]], 0).to(bf).contiguous()
    T["bv"] = torch.cat([lw["ca_in_b"][2 * E:] for lw in P["layers"]], 0).contiguous()
    P["tc"] = T
    return T


def key_side_tc(eng, x, img_metas, taps=None):
    """Everything that depends on the image tokens: masks, input_proj, position embedding,
    K/V projection of all layers.  Returns the key-side state: Kh, Vh (L,B,heads,Lk,32) bf16,
    kmask (None when no key is padded), the split-KV plan."""
    s = eng.spec
    x = x.float().contiguous()
    B, N, Cin, H, W = x.shape
    dev = x.device
    Lk = N * H * W
    P = eng.packed(dev)
    T = packed_tc(eng, dev)
    masks = eng.masks(img_metas, B, N, H, W, dev)
    has_mask = any_key_masked(img_metas, N, H, W)
    vin = torch.empty((B * Lk, s.E), dtype=torch.bfloat16, device=dev)
    mem = ops.gemm_f32(x, P["in_w"], P["in_b"], nchw_hw=H * W, out_bf16=vin)
    sf = ops.sine3d_features(masks, P["dim_t_sine"], dtype=torch.bfloat16, **_sine_args(s))
    h2 = ops.gemm_bf16_tc(sf, T["ad_w1"], P["ad_b1"], act=ACT_RELU, out_dtype=torch.bfloat16)
    if s.with_position:
        feat, _, _ = eng.frustum(img_metas, B, N, H, W, dev, None, torch.bfloat16, want_mask=False)
        h1 = ops.gemm_bf16_tc(feat, T["pe_w1"], P["pe_b1"], act=ACT_RELU, out_dtype=torch.bfloat16)
        if s.with_fpe:
            pos = ops.gemm_bf16_tc(h1, T["pe_w2"], P["pe_b2"])
            se = ops.gemm_f32(mem, P["fpe_wr"], P["fpe_br"], act=ACT_RELU)
            pos = ops.gemm_f32(se, P["fpe_we"], P["fpe_be"], act=ACT_SIGMOID_GATE, gate=pos, residual=mem)
        else:
            pos = ops.gemm_bf16_tc(h1, T["pe_w2"], P["pe_b2"], residual=mem)      # mem + position MLP
    else:
        pos = mem
    kin = ops.gemm_bf16_tc(h2, T["ad_w2"], P["ad_b2"], residual=pos, out_dtype=torch.bfloat16)
    Kh = ops.gemm_bf16_tc(kin, T["wk"], T["bk"], headsplit_Lk=Lk)
    Vh = ops.gemm_bf16_tc(vin, T["wv"], T["bv"], headsplit_Lk=Lk)
    if taps is not None:
        taps.update(masks=masks, memory=mem, kin=kin, Kh=Kh, Vh=Vh)
    kmask = masks.view(B, Lk) if has_mask else None
    if s.E != 256 or s.heads != 8:
        raise NotImplementedError("tcgen05 path is specialised for embed_dims 256, 8 heads of 32")
    nsplit = choose_nsplit(B, s.heads, s.Lq, Lk)
    partial = None
    if nsplit > 1:
        partial = torch.empty((nsplit * ops.attention_partial_floats(B, s.Lq, s.heads),),
                              dtype=torch.float32, device=dev)
    return dict(precision="bf16", B=B, Lk=Lk, Kh=Kh, Vh=Vh, kmask=kmask, nsplit=nsplit, partial=partial)


def _sine_args(s):
    a = s.sine
    return dict(num_feats=a["num_feats"], normalize=a["normalize"], scale=a["scale"],
                eps=a["eps"], offset=a["offset"])


def cross_attn_tc(eng, X, qpos, lw, l: int, ks: dict):
    """Cross-attention block of decoder layer l on the tcgen05 path: fp32 Q projection ->
    bf16 pack (scale folded) -> flash attention over the pre-projected K/V -> fp32 out-proj +
    residual -> LayerNorm (petr_transformer.py:341-367)."""
    s = eng.spec
    E = s.E
    B, Lk = ks["B"], ks["Lk"]
    q = ops.gemm_f32(X, lw["ca_in_w"][:E], lw["ca_in_b"][:E], A2=qpos)
    qs = ops.pack_q_bf16(q, B, s.Lq, s.heads)
    o = ops.cross_attention_bf16(qs, ks["Kh"][l], ks["Vh"][l], ks["kmask"], B, s.Lq, Lk, s.heads,
                                 ks["nsplit"], partial=ks["partial"])
    y = ops.gemm_f32(o, lw["ca_out_w"], lw["ca_out_b"], residual=X)
    return ops.layernorm(y, *lw["ln"][1], out=y)
