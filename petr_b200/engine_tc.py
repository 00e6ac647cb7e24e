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
DEDUP_SINE = True                # evaluate the batch-invariant sine/adapt embedding once per forward
CACHE_CONST_SINE = True          # no padded key anywhere: the sine / adapt embedding is a constant of (N, H, W) and the weights
SPLIT_K_SMALL_M = True           # split-K for the query-side residual GEMMs when B * 900 rows leave most SMs idle
FOLD_KV_BIAS = True              # drop the K bias (softmax-invariant), fold the V bias into the out-projection bias
DEDUP_LAYER0 = True              # layer 0's self-attention block sees X = 0 for every sample: evaluate it for one
FUSED_POSITION_MLP = True        # False: unfused frustum kernel + two GEMMs (kept for A/B tests)


from .engine import MetaInputs, any_key_masked  # noqa: F401  (re-exported for tests)


def choose_nsplit(B: int, heads: int, Lq: int, Lk: int) -> int:
    """Key splits so that the grid fills the 148 SMs in whole waves (1 CTA per SM; a CTA of the
    flash kernel owns two 128-query tiles of one (batch, head))."""
    base = B * heads * ((Lq + 255) // 256)
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


def choose_ksplit(rows: int, k_blocks: int) -> int:
    """Split-K parts for a query-side GEMM with a 256-wide output (one column of 128-row tiles): at small batch the
    B * 900 rows are a handful of tiles (8 at B = 1) and most SMs idle while each tile walks its whole K range.  The
    parts are summed by the LayerNorm that follows (fixed order: deterministic)."""
    tiles = (rows + 127) // 128
    if not SPLIT_K_SMALL_M or tiles * 2 > NUM_SMS:
        return 1
    return max(1, min(NUM_SMS // tiles, k_blocks // 4, 8))


def packed_tc(eng, device) -> dict:
    """bf16 weight copies for the tensor-core path: plain bf16 for the Lk-side GEMMs, [hi|hi|lo]
    (ops.x3_weight) for the fp32-class query-side GEMMs."""
    P = eng.packed(device)
    if "tc" in P:
        return P["tc"]
    s = eng.spec
    sd = P["sd"]
    E = s.E
    bf = torch.bfloat16
    x3 = ops.x3_weight
    T = dict()
    T["in_w"] = P["in_w"].to(bf).contiguous()
    if s.with_position:
        T["pe_w1"] = P["pe_w1"].to(bf).contiguous()
        T["pe_w2"] = P["pe_w2"].to(bf).contiguous()
    if s.with_fpe:
        T["fpe_wr"] = x3(P["fpe_wr"])
        T["fpe_we"] = x3(P["fpe_we"])
    T["ad_w1"] = P["ad_w1"].to(bf).contiguous()
    T["ad_w2"] = P["ad_w2"].to(bf).contiguous()
    # K / V projection of all layers as two GEMMs WITHOUT bias (FOLD_KV_BIAS): the K bias adds q.bk to every score of a
    # query row -- constant over the keys, so the softmax does not see it; the V bias passes through the attention
    # unchanged (the probabilities sum to one) and is folded into the out-projection bias, bo' = bo + Wo bv.  Exact in
    # real arithmetic; it takes the bias tile, its barrier and 32 adds per chunk out of the store-bound epilogue.
    T["wk"] = torch.cat([lw["ca_in_w"][E:2 * E] for lw in P["layers"]], 0).to(bf).contiguous()
    T["wv"] = torch.cat([lw["ca_in_w"][2 * E:] for lw in P["layers"]], 0).to(bf).contiguous()
    if FOLD_KV_BIAS:
        T["bk"] = T["bv"] = None
    else:
        T["bk"] = torch.cat([lw["ca_in_b"][E:2 * E] for lw in P["layers"]], 0).contiguous()
        T["bv"] = torch.cat([lw["ca_in_b"][2 * E:] for lw in P["layers"]], 0).contiguous()
    T["layers"] = []
    for lw in P["layers"]:
        T["layers"].append(dict(
            sa_qk=x3(lw["sa_in_w"][:2 * E]), sa_qk_b=lw["sa_in_b"][:2 * E].contiguous(),
            sa_v=x3(lw["sa_in_w"][2 * E:]), sa_v_b=lw["sa_in_b"][2 * E:].contiguous(),
            sa_out=x3(lw["sa_out_w"]), ca_q=x3(lw["ca_in_w"][:E]), ca_q_b=lw["ca_in_b"][:E].contiguous(),
            ca_out=x3(lw["ca_out_w"]), f1=x3(lw["f1_w"]), f2=x3(lw["f2_w"]),
            ca_out_b=((lw["ca_out_b"].double() + lw["ca_out_w"].double() @ lw["ca_in_b"][2 * E:].double()).float().contiguous()
                      if FOLD_KV_BIAS else lw["ca_out_b"])))
    nb = 1 if s.shared_branches else s.L
    T["branches"] = []
    for l in range(nb):
        c, r = f"cls_branches.{l}.", f"reg_branches.{l}."
        d = dict(cls=[x3(sd[f"{c}{3 * i}.weight"]) for i in range(s.num_reg_fcs)])
        if s.with_multi:
            d["reg"] = [x3(sd[f"{r}reg_branch.{3 * i}.weight"]) for i in range(s.num_reg_fcs)]
            d["task0"] = x3(P[f"task_w0_{l}"])
        else:
            d["reg"] = [x3(sd[f"{r}{2 * i}.weight"]) for i in range(s.num_reg_fcs)]
        T["branches"].append(d)
    P["tc"] = T
    return T


def slice_img_metas(img_metas, views: slice):
    """img_metas restricted to a contiguous range of camera views (key-split over views)."""
    out = []
    for m in img_metas:
        d = dict(m)
        for k in ("lidar2img", "img_shape"):
            d[k] = list(m[k])[views]
        out.append(d)
    return out


def shard_views(num_views: int, world: int, rank: int) -> slice:
    """Contiguous, balanced split of the camera views over ranks (rank r owns views [lo, hi))."""
    base, rem = divmod(num_views, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def key_side_tc(eng, x, mi: MetaInputs, taps=None, views: slice = None):
    """Everything that depends on the image tokens: masks, input_proj, position embedding,
    K/V projection of all layers.  Returns the key-side state: Kh, Vh (L,B,heads,Lk,32) bf16,
    kmask (None when no key is padded), the split-KV plan.

    views: this rank's contiguous range of camera views (key-split).  The
    sine encoding is cumulative over ALL views (positional_encoding.py:71), so masks and sine
    features are evaluated for every view (cheap) and only the local rows are kept."""
    s = eng.spec
    if s.E != 256 or s.heads != 8:
        raise NotImplementedError("tcgen05 path is specialised for embed_dims 256, 8 heads of 32")
    if x.dtype != torch.bfloat16:            # bf16 feature maps are taken as they are (half the bytes on the way in)
        x = x.float()
    B, N_all, Cin, H, W = x.shape
    if Cin % 64 != 0:
        raise NotImplementedError("tcgen05 path needs in_channels % 64 == 0")
    dev = x.device
    E = s.E
    P = eng.packed(dev)
    T = packed_tc(eng, dev)
    masks_all = eng.masks(mi, H, W)
    if views is not None:
        x = x[:, views]
        masks = masks_all[:, views].contiguous()
        mi = mi.views(views)
    else:
        masks = masks_all
    x = x.contiguous()
    N = x.shape[1]
    Lk = N * H * W
    has_mask = mi.has_mask
    # Sine encoding + adapt_pos3d (petr_head.py:399-402).  It depends on the key masks only; when every
    # sample of the batch has the same img_shape list (all shipped configs: no padding at all) the
    # masks are identical per sample, so the (Lk, 256) result is evaluated once per forward for one
    # sample and broadcast over the batch instead of B times.
    dedup = DEDUP_SINE and mi.uniform_masks
    m_sine = masks_all[:1] if dedup else masks_all
    raw_pos = not s.with_position and not s.with_multiview      # petr_head.py:414-420: encoding used as is
    # With no padded key at all (every shipped config: final_dim divisible by 32, img_shape == pad_shape; known on the
    # host) the encoding is a constant of (N, H, W) and the adapt_pos3d weights: SURVEY 8a3 "cache per (N,H,W)".  It then
    # lives in the packed state (dropped with the weights); tensors made during graph capture are not kept.
    ckey = ("sine_const", N_all, H, W, None if views is None else (views.start, views.stop), raw_pos)
    const = CACHE_CONST_SINE and dedup and not has_mask
    if const and ckey in T:
        h2, sin1 = None, T[ckey]
    else:
        sf = eng.pos_features(P, m_sine, torch.float32 if raw_pos else torch.bfloat16)
        if views is not None:
            sf = sf.view(m_sine.shape[0], N_all, H * W, -1)[:, views].reshape(m_sine.shape[0] * Lk, -1).contiguous()
        if raw_pos:
            h2, sin1 = None, sf                                                          # (Lk | B*Lk, E) fp32
        else:
            h2 = ops.gemm_bf16_tc(sf, T["ad_w1"], P["ad_b1"], act=ACT_RELU, out_dtype=torch.bfloat16, name="gemm_adapt1")
            sin1 = ops.gemm_bf16_tc(h2, T["ad_w2"], P["ad_b2"], name="gemm_adapt2") if dedup else None   # (Lk, E) fp32
        if const and sin1 is not None and not torch.cuda.is_current_stream_capturing():
            T[ckey] = sin1
    if raw_pos:
        dedup = True                                     # add_rows_bf16 below takes either period
    fused = s.with_position and FUSED_POSITION_MLP and s.D == 64 and tuple(T["pe_w1"].shape) == (1024, 192) and \
        tuple(T["pe_w2"].shape) == (256, 1024)
    # input_proj (petr_head.py:390): fp32 NCHW -> bf16 tokens -> tcgen05 GEMM.  One accumulator, two results
    # (out_mode 5): the bf16 tokens that feed the V projection, and fp32 "memory + broadcast sine embedding", the
    # single residual of the fused position MLP whose bf16 output is then directly the K-projection input.
    xt = ops.nchw_to_tokens_bf16(x.view(B * N, Cin, H, W))
    mem = torch.empty((B * Lk, E), dtype=torch.float32, device=dev)
    fold_sine = dedup and fused and not s.with_fpe and not raw_pos and sin1.shape[0] == Lk
    if fold_sine:
        vin = ops.gemm_bf16_tc(xt, T["in_w"], P["in_b"], residual=sin1, residual_period=Lk, out_f32=mem,
                               name="gemm_input_proj")
    else:
        vin_split = ops.gemm_bf16_tc(xt, T["in_w"], P["in_b"], split_out=True, out_f32=mem, name="gemm_input_proj")
        vin = vin_split[:, :E]
    kin = None
    if s.with_position:
        res = None if s.with_fpe else mem                       # mem + position MLP in the epilogue
        if fused:
            # K1: frustum -> inverse_sigmoid -> MLP in one tcgen05 kernel, nothing intermediate in HBM
            if fold_sine:
                kin = ops.position_mlp_fused(mi.i2l, H, W, s.D, mi.pad_h, mi.pad_w, s.lid, s.depth_start,
                                             s.bin_size, s.position_range, T["pe_w1"], P["pe_b1"], T["pe_w2"],
                                             P["pe_b2"], residual=mem, want_f32=False,
                                             out_bf16=torch.empty((B * Lk, E), dtype=torch.bfloat16, device=dev))
            else:
                pos = ops.position_mlp_fused(mi.i2l, H, W, s.D, mi.pad_h, mi.pad_w, s.lid, s.depth_start,
                                             s.bin_size, s.position_range, T["pe_w1"], P["pe_b1"], T["pe_w2"],
                                             P["pe_b2"], residual=res)
        else:
            feat, _, _ = eng.frustum(mi, B, N, H, W, None, torch.bfloat16, want_mask=False)
            h1 = ops.gemm_bf16_tc(feat, T["pe_w1"], P["pe_b1"], act=ACT_RELU, out_dtype=torch.bfloat16, name="gemm_pe1")
            pos = ops.gemm_bf16_tc(h1, T["pe_w2"], P["pe_b2"], residual=res, name="gemm_pe2")
        if s.with_fpe:
            # SELayer (petrv2_head.py:48-60,:464-466) as two bf16x3 tcgen05 GEMMs on the [hi | lo] tokens input_proj
            # already emitted; the gate x position embedding product and the "+ memory" are the second epilogue
            se = ops.gemm_bf16_tc(vin_split, T["fpe_wr"], P["fpe_br"], act=ACT_RELU, x3=True, split_out=True,
                                  name="gemm_fpe")
            pos = ops.gemm_bf16_tc(se, T["fpe_we"], P["fpe_be"], act=ACT_SIGMOID_GATE, gate=pos, residual=mem, x3=True,
                                   name="gemm_fpe")
    else:
        pos = mem
    if kin is None:
        if dedup:
            kin = ops.add_rows_bf16(pos, sin1)
        else:
            kin = ops.gemm_bf16_tc(h2, T["ad_w2"], P["ad_b2"], residual=pos, out_dtype=torch.bfloat16, name="gemm_adapt2")
    Kh = ops.gemm_bf16_tc(kin, T["wk"], T["bk"], headsplit_Lk=Lk, name="gemm_kproj")
    Vh = ops.gemm_bf16_tc(vin, T["wv"], T["bv"], headsplit_Lk=Lk, name="gemm_vproj")
    if taps is not None:
        taps.update(masks=masks, memory=None if fold_sine else mem, kin=kin, Kh=Kh, Vh=Vh)
    kmask = masks.view(B, Lk) if has_mask else None
    nsplit = choose_nsplit(B, s.heads, s.Lq, Lk)
    partial = None
    if nsplit > 1:
        partial = torch.empty((nsplit * ops.attention_partial_floats(B, s.Lq, s.heads),),
                              dtype=torch.float32, device=dev)
    att_ws = ops.cross_attention_workspace(B, s.Lq, Lk, s.heads, nsplit, dev)
    return dict(precision="bf16", B=B, Lk=Lk, Kh=Kh, Vh=Vh, kmask=kmask, nsplit=nsplit, partial=partial,
                att_ws=att_ws)


def _sine_args(s):
    a = s.sine
    return dict(num_feats=a["num_feats"], normalize=a["normalize"], scale=a["scale"],
                eps=a["eps"], offset=a["offset"])


Q_SCALE = (1.0 / 32.0 ** 0.5) * 1.4426950408889634          # head_dim^-1/2 * log2(e)


def decoder_layer_tc(eng, P, l: int, st: dict, qpos, ks: dict, last: bool) -> dict:
    """One decoder layer on the tensor-core path.  `st` carries the residual stream in the forms
    the next GEMMs need: X fp32 (R,E), Xs = split(X), Xps = split(X + query_pos) (bf16 [hi|lo]).
    Every projection / FFN GEMM is a bf16x3 tcgen05 GEMM (fp32-class accuracy); the self-attention
    core is the x3 (hi/lo split) tcgen05 kernel; the cross-attention core is the bf16 flash kernel."""
    s = eng.spec
    E, R, B = s.E, st["X"].shape[0], ks["B"]
    lw = P["layers"][l]
    tw = packed_tc(eng, st["X"].device)["layers"][l]
    dev = st["X"].device
    g = ops.gemm_bf16_tc
    o_s = torch.empty((R, 2 * E), dtype=torch.bfloat16, device=dev)
    # The decoder starts from target = zeros for every sample (petr_transformer.py:97): the self-attention block of the
    # first layer, its LayerNorm and the cross-attention's Q projection see the same input in every sample of the
    # batch.  Evaluated once per forward for one sample (the same kernels on 900 rows) and broadcast -- bit-identical to
    # evaluating it B times (rows are independent in the GEMMs / LayerNorm, the attention CTAs are per sample).
    n1 = 1 if (DEDUP_LAYER0 and st.get("batch_invariant") and B > 1) else B
    r1 = n1 * s.Lq
    # ---- self-attention (mmcv MultiheadAttention; q = k = X + qpos, v = X) ----
    qk = g(st["Xps"][:r1], tw["sa_qk"], tw["sa_qk_b"], x3=True, headsplit_Lk=s.Lq, headsplit_pair=True,
           name="gemm_q_side")                               # (hi/lo, [Q, K], B, heads, Lq, 32)
    v = g(st["Xs"][:r1], tw["sa_v"], tw["sa_v_b"], x3=True, headsplit_Lk=s.Lq, headsplit_pair=True,
          name="gemm_q_side")
    ops.attention_x3_tc(qk[:, 0], qk[:, 1], v[:, 0], None, n1, s.Lq, s.Lq, s.heads, out_split=o_s[:r1],
                        want_f32=False)
    kp = choose_ksplit(r1, 3 * E // 64)
    y = g(o_s[:r1], tw["sa_out"], lw["sa_out_b"], residual=st["X"][:r1], x3=True, ksplit=kp, name="gemm_q_side")
    X1 = torch.empty((r1, E), dtype=torch.float32, device=dev)
    X1ps = torch.empty((r1, 2 * E), dtype=torch.bfloat16, device=dev)
    ops.layernorm_multi(y, *lw["ln"][0], qpos=qpos, Y=X1, Ypos_split=X1ps, nparts=kp)
    # ---- cross-attention (PETRMultiheadAttention) ----
    qs = g(X1ps, tw["ca_q"], tw["ca_q_b"], x3=True, alpha=Q_SCALE, headsplit_Lk=s.Lq, name="gemm_q_side")
    if n1 != B:
        X1 = X1.repeat(B, 1)
        qs = qs.expand(qs.shape[0], B, *qs.shape[2:]).contiguous()
    if "xattn" in ks:
        ks["xattn"](qs[0], l, o_s)                           # key-split: partials + exchange + merge
    else:
        ops.cross_attention_bf16(qs[0], ks["Kh"][l], ks["Vh"][l], ks["kmask"], B, s.Lq, ks["Lk"], s.heads,
                                 ks["nsplit"], partial=ks["partial"], out_split=o_s, want_f32=False,
                                 workspace=ks["att_ws"])
    kp = choose_ksplit(R, 3 * E // 64)
    y = g(o_s, tw["ca_out"], tw["ca_out_b"], residual=X1, x3=True, ksplit=kp, name="gemm_q_side")
    X2 = torch.empty((R, E), dtype=torch.float32, device=dev)
    X2s = torch.empty((R, 2 * E), dtype=torch.bfloat16, device=dev)
    ops.layernorm_multi(y, *lw["ln"][1], Y=X2, Y_split=X2s, nparts=kp)
    # ---- FFN ----
    hs = g(X2s, tw["f1"], lw["f1_b"], act=ACT_RELU, x3=True, split_out=True, name="gemm_q_side")
    kp = choose_ksplit(R, 3 * hs.shape[1] // 2 // 64)
    y = g(hs, tw["f2"], lw["f2_b"], residual=X2, x3=True, ksplit=kp, name="gemm_q_side")
    X3 = torch.empty((R, E), dtype=torch.float32, device=dev)
    X3s = torch.empty((R, 2 * E), dtype=torch.bfloat16, device=dev)
    X3ps = None if last else torch.empty((R, 2 * E), dtype=torch.bfloat16, device=dev)
    ops.layernorm_multi(y, *lw["ln"][2], qpos=None if last else qpos, Y=X3, Y_split=X3s, Ypos_split=X3ps,
                        gamma2=P["post_norm"][0], beta2=P["post_norm"][1], Z=st["outs"][l],
                        Z_split=st["outs_split"][l], nparts=kp)
    return dict(X=X3, Xs=X3s, Xps=X3ps, outs=st["outs"], outs_split=st["outs_split"])


def branches_tc(eng, P, outs, outs_split, B: int, time_div=None):
    """cls / reg branches on the tensor-core path: the 256x256 layers are bf16x3 tcgen05 GEMMs on
    the split post_norm outputs, the final 256 -> 10 layers and the box decode stay fp32 kernels
    (petr_head.py:440-460)."""
    s = eng.spec
    sd = P["sd"]
    L = outs.shape[0]
    R = B * s.Lq
    E = s.E
    T = packed_tc(eng, outs.device)
    g = ops.gemm_bf16_tc
    dev = outs.device
    cls = torch.empty((L * R, s.cls_out), dtype=torch.float32, device=dev)
    reg = torch.empty((L * R, s.code_size), dtype=torch.float32, device=dev)
    groups = [(0, outs_split.view(L * R, 2 * E), slice(0, L * R))] if s.shared_branches else \
        [(l, outs_split[l], slice(l * R, (l + 1) * R)) for l in range(L)]
    for l, xs, rows in groups:
        bw = T["branches"][0 if s.shared_branches else l]
        c, r = f"cls_branches.{l}.", f"reg_branches.{l}."
        n = xs.shape[0]
        h = xs
        yf = None
        for i in range(s.num_reg_fcs):
            y = g(h, bw["cls"][i], sd[f"{c}{3 * i}.bias"], x3=True, name="gemm_q_side")
            h = torch.empty((n, 2 * E), dtype=torch.bfloat16, device=dev)
            yf = torch.empty((n, E), dtype=torch.float32, device=dev)
            ops.layernorm_multi(y, sd[f"{c}{3 * i + 1}.weight"], sd[f"{c}{3 * i + 1}.bias"], Y=yf, Y_split=h, relu=True)
        j = 3 * s.num_reg_fcs
        eng.cls_last(sd, c, yf, B, out=cls[rows])
        if s.with_multi:
            h = xs
            for i in range(s.num_reg_fcs):
                h = g(h, bw["reg"][i], sd[f"{r}reg_branch.{3 * i}.bias"], act=ACT_RELU, x3=True, split_out=True,
                      name="gemm_q_side")
            t = g(h, bw["task0"], P[f"task_b0_{l}"], act=ACT_RELU, x3=True, name="gemm_q_side")      # (n, 5E) fp32
            c0 = 0
            for k in range(5):
                w = sd[f"{r}task_heads.{k}.2.weight"]
                ops.gemm_f32(t[:, k * E:(k + 1) * E], w, sd[f"{r}task_heads.{k}.2.bias"],
                             out=reg[rows][:, c0:c0 + w.shape[0]])
                c0 += w.shape[0]
        else:
            h = xs
            for i in range(s.num_reg_fcs):
                lastfc = i == s.num_reg_fcs - 1
                if lastfc:
                    h = g(h, bw["reg"][i], sd[f"{r}{2 * i}.bias"], act=ACT_RELU, x3=True, name="gemm_q_side")
                else:
                    h = g(h, bw["reg"][i], sd[f"{r}{2 * i}.bias"], act=ACT_RELU, x3=True, split_out=True,
                          name="gemm_q_side")
            j = 2 * s.num_reg_fcs
            ops.gemm_f32(h, sd[f"{r}{j}.weight"], sd[f"{r}{j}.bias"], out=reg[rows])
    ops.decode_boxes(reg, P["ref"], s.pc_range, time_div, L, B, s.Lq)
    return cls.view(L, B, s.Lq, s.cls_out), reg.view(L, B, s.Lq, s.code_size)


def initial_state(eng, B: int, qpos, dev) -> dict:
    s = eng.spec
    R = B * s.Lq
    X = torch.zeros((R, s.E), dtype=torch.float32, device=dev)
    Xs = torch.zeros((R, 2 * s.E), dtype=torch.bfloat16, device=dev)
    Xps = ops.split_rows(X, add=qpos)
    outs = torch.empty((s.L, R, s.E), dtype=torch.float32, device=dev)
    outs_split = torch.empty((s.L, R, 2 * s.E), dtype=torch.bfloat16, device=dev)
    return dict(X=X, Xs=Xs, Xps=Xps, outs=outs, outs_split=outs_split, batch_invariant=True)


def state_from_rows(eng, X, qpos, st_like: dict) -> dict:
    """Teacher-forcing helper (tests): rebuild the split forms of an externally supplied X."""
    return dict(X=X, Xs=ops.split_rows(X), Xps=ops.split_rows(X, add=qpos), outs=st_like["outs"],
                outs_split=st_like["outs_split"])


# ----------------------------------------------------------------------------------------------
# key-split over camera views: rank r owns a contiguous range of views, runs
# the key side and the partial (un-normalised) cross-attention on its keys; per layer the packed
# [O | m, l] partials are exchanged with ONE all-gather and merged by petr_attention_merge (the same
# kernel and layout as on-GPU split-KV).  Queries / self-attention / FFN / branches are replicated.
# ----------------------------------------------------------------------------------------------
def keysplit_nsplit(s, B: int, x, world: int) -> int:
    """Split-KV factor INSIDE every rank of a key-split (each rank still has to fill its 148 SMs: at B = 1 one
    (head, query-tile pair) grid is 32 CTAs).  The same on all ranks -- derived from the smallest shard -- because the
    gathered buffer is laid out [rank][split]."""
    N, H, W = x.shape[1], x.shape[3], x.shape[4]
    views = min(shard_views(N, world, r).stop - shard_views(N, world, r).start for r in range(world))
    return choose_nsplit(B, s.heads, s.Lq, views * H * W)


def key_side_keysplit(eng, x, mi: MetaInputs, world: int, rank: int, group=None, taps=None) -> dict:
    """Key-side state of one rank of a `world`-way key-split (torch.distributed NCCL group)."""
    import torch.distributed as dist
    s = eng.spec
    N = x.shape[1]
    ks = key_side_tc(eng, x, mi, taps, views=shard_views(N, world, rank))
    B = ks["B"]
    ns = keysplit_nsplit(s, B, x, world)
    ws = ops.cross_attention_workspace(B, s.Lq, ks["Lk"], s.heads, ns, x.device)
    floats = ns * ops.attention_partial_floats(B, s.Lq, s.heads)
    local = torch.empty((floats,), dtype=torch.float32, device=x.device)
    gathered = torch.empty((world * floats,), dtype=torch.float32, device=x.device)

    def xattn(qs, l, o_s):
        ops.cross_attention_bf16(qs, ks["Kh"][l], ks["Vh"][l], ks["kmask"], B, s.Lq, ks["Lk"], s.heads, ns,
                                 partial=local, write_partial=True, workspace=ws)
        dist.all_gather_into_tensor(gathered, local, group=group)
        ops.attention_merge(gathered, world * ns, B, s.Lq, s.heads, out_split=o_s, want_f32=False)

    ks["xattn"] = xattn
    return ks


class P2PExchange:
    """Peer-memory exchange state of one rank of a key-split (include/petr_b200.h, "Key-split exchange over NVLink peer
    memory"): one symmetric allocation (torch.distributed._symmetric_memory: cuMem + IPC handles, no NVSHMEM) holding
    gathered[2][world][floats] fp32 and flags[2][world] uint32, plus the device arrays of peer pointers the push
    kernel dereferences.  Creation is collective (rendezvous + barrier) and must happen outside graph capture."""

    def __init__(self, world: int, rank: int, floats: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank, self.floats = world, rank, floats
        n = 2 * world * floats
        self.buf = symm_mem.empty(n + 64, dtype=torch.float32, device=device)
        self.buf.zero_()
        pg = group if group is not None else dist.group.WORLD
        try:                                      # needed by some torch versions, deprecated no-op in others
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                symm_mem.enable_symm_mem_for_group(pg.group_name)
        except Exception:
            pass
        hdl = symm_mem.rendezvous(self.buf, pg)
        self.hdl = hdl
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.peer_bufs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.peer_flags = torch.tensor([p + n * 4 for p in ptrs], dtype=torch.int64, device=device)
        self.flags = self.buf[n:n + 2 * world].view(torch.int32)
        self.epoch = torch.zeros((1,), dtype=torch.int32, device=device)
        self.counters = torch.zeros((world,), dtype=torch.int32, device=device)
        self.err = torch.zeros((1,), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                       # nobody pushes before every rank has zeroed its flags

    def slot(self, parity: int, r: int) -> torch.Tensor:
        o = (parity * self.world + r) * self.floats
        return self.buf[o:o + self.floats]

    def gathered(self, parity: int) -> torch.Tensor:
        o = parity * self.world * self.floats
        return self.buf[o:o + self.world * self.floats]


def key_side_keysplit_p2p(eng, x, mi: MetaInputs, world: int, rank: int, group=None, taps=None) -> dict:
    """Key-side state of one rank of a `world`-way key-split whose per-layer exchange is a push over NVLink peer memory
    (petr_p2p_push_partial / petr_attention_merge_p2p) instead of an NCCL all-gather: graph-capturable."""
    from ._lib import call, ptr, stream_ptr
    s = eng.spec
    N = x.shape[1]
    ks = key_side_tc(eng, x, mi, taps, views=shard_views(N, world, rank))
    B = ks["B"]
    ns = keysplit_nsplit(s, B, x, world)
    ws = ops.cross_attention_workspace(B, s.Lq, ks["Lk"], s.heads, ns, x.device)
    floats = ns * ops.attention_partial_floats(B, s.Lq, s.heads)
    key = ("p2p", world, rank, floats, str(x.device))
    st = eng._p2p.get(key) if hasattr(eng, "_p2p") else None
    if st is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("petr_b200: the peer-memory exchange must be set up by an un-captured forward first")
        if not hasattr(eng, "_p2p"):
            eng._p2p = {}
        st = eng._p2p[key] = P2PExchange(world, rank, floats, x.device, group)
    dev = x.device
    call("petr_p2p_epoch_inc", ptr(st.epoch), stream_ptr(dev))

    def xattn(qs, l, o_s):
        par = l & 1
        ops.cross_attention_bf16(qs, ks["Kh"][l], ks["Vh"][l], ks["kmask"], B, s.Lq, ks["Lk"], s.heads, ns,
                                 partial=st.slot(par, rank), write_partial=True, workspace=ws)
        call("petr_p2p_push_partial", ptr(st.slot(par, rank)), ptr(st.peer_bufs), ptr(st.peer_flags), world, rank, floats,
             par, ptr(st.epoch), l, s.L, ptr(st.counters), stream_ptr(dev))
        call("petr_attention_merge_p2p", ptr(st.gathered(par)), world * ns, world, ptr(st.flags[par * world:]), ptr(st.epoch), l, s.L,
             ptr(st.err), None, 0, ptr(o_s), B, s.Lq, s.heads, stream_ptr(dev))

    ks["xattn"] = xattn
    ks["p2p"] = st
    return ks


def key_side_keysplit_emulated(eng, x, mi: MetaInputs, world: int, taps=None) -> dict:
    """Single-GPU emulation of a `world`-way key-split: every rank's shard is processed in turn on
    this device and the partials are laid out exactly as the all-gather would (tests; also the
    documented way to exercise the N-rank path with fewer GPUs than ranks)."""
    s = eng.spec
    N = x.shape[1]
    shards = [key_side_tc(eng, x, mi, taps if r == 0 else None, views=shard_views(N, world, r))
              for r in range(world)]
    B = shards[0]["B"]
    ns = keysplit_nsplit(s, B, x, world)
    wss = [ops.cross_attention_workspace(B, s.Lq, k["Lk"], s.heads, ns, x.device) for k in shards]
    floats = ns * ops.attention_partial_floats(B, s.Lq, s.heads)
    gathered = torch.empty((world * floats,), dtype=torch.float32, device=x.device)

    def xattn(qs, l, o_s):
        for r, k in enumerate(shards):
            ops.cross_attention_bf16(qs, k["Kh"][l], k["Vh"][l], k["kmask"], B, s.Lq, k["Lk"], s.heads, ns,
                                     partial=gathered[r * floats:(r + 1) * floats], write_partial=True,
                                     workspace=wss[r])
        ops.attention_merge(gathered, world * ns, B, s.Lq, s.heads, out_split=o_s, want_f32=False)

    out = dict(shards[0])
    out["xattn"] = xattn
    out["Lk"] = sum(k["Lk"] for k in shards)
    return out
