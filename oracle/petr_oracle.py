"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference PETR head hot path.

This is the *oracle*: a functional, state_dict-driven restatement (torch CPU fp32 + numpy
float64 on the host, exactly where the reference uses them) of

    PETRHead.forward / position_embeding      projects/mmdet3d_plugin/models/dense_heads/petr_head.py:286-468
    PETRv2Head.forward (SELayer, RegLayer)    .../dense_heads/petrv2_head.py:48-95, :348-540
    PETRTransformer / Decoder / DecoderLayer  .../models/utils/petr_transformer.py:69-109, :423-447
    PETRMultiheadAttention                    .../models/utils/petr_transformer.py:280-367
    SinePositionalEncoding3D                  .../models/utils/positional_encoding.py:58-100
    mmcv BaseTransformerLayer / FFN / MultiheadAttention (un-vendored mmcv-full 1.4.0; the
        reference pins them through its vendored copies multi_atten_decoder_layer.py:72-293
        and petr_transformer.py:228-367)
    mmdet inverse_sigmoid (vendored copy utils/detr.py:15-31)

Pinning: ``tests/test_oracle_vs_reference.py`` runs this file against the UNMODIFIED reference
(loaded by ``oracle/ref_loader.py``) in the build container, and ``tests/golden/*.npz`` hold
outputs of the reference itself (``oracle/make_golden.py``) which the oracle must reproduce
everywhere, including on the GPU box where /root/reference does not exist.  The reference
ships no tests or golden vectors of its own, so these are the pins.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product (``petr_b200``) never does.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------
def inverse_sigmoid(x: Tensor, eps: float = 1e-5) -> Tensor:
    """mmdet inverse_sigmoid; reference copy at models/utils/detr.py:15-31."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def frustum_axes(H: int, W: int, pad_h: int, pad_w: int, depth_num: int = 64,
                 depth_start: float = 1.0, position_range=(-61.2, -61.2, -10.0, 61.2, 61.2, 10.0),
                 LID: bool = True, depth_step: float = 0.8):
    """petr_head.py:290-302.  fp32 pixel-corner grids and depth bins (bit-exact class)."""
    coords_h = torch.arange(H).float() * pad_h / H
    coords_w = torch.arange(W).float() * pad_w / W
    index = torch.arange(0, depth_num, 1).float()
    if LID:
        bin_size = (position_range[3] - depth_start) / (depth_num * (1 + depth_num))
        coords_d = depth_start + bin_size * index * (index + 1)
    else:
        bin_size = (position_range[3] - depth_start) / depth_num
        coords_d = depth_start + bin_size * index
    return coords_h, coords_w, coords_d


def key_padding_masks(img_metas: Sequence[dict], B: int, N: int, H: int, W: int) -> Tensor:
    """petr_head.py:383-394: ones(pad) with zeros over each view's img_shape, nearest-resized
    to the feature map.  Returns bool (B,N,H,W), True = padded key."""
    pad_h, pad_w, _ = img_metas[0]["pad_shape"][0]
    masks = torch.ones((B, N, pad_h, pad_w), dtype=torch.float32)
    for b in range(B):
        for n in range(N):
            img_h, img_w, _ = img_metas[b]["img_shape"][n]
            masks[b, n, :img_h, :img_w] = 0
    return F.interpolate(masks, size=(H, W)).to(torch.bool)


def img2lidar_matrices(img_metas: Sequence[dict]) -> Tensor:
    """petr_head.py:308-315: float64 np.linalg.inv per view on the host, then one rounding
    to fp32.  Returns (B,N,4,4) fp32."""
    out = []
    for meta in img_metas:
        out.append(np.asarray([np.linalg.inv(m) for m in meta["lidar2img"]]))
    return torch.tensor(np.asarray(out), dtype=torch.float32)


def frustum_coords3d(img2lidars: Tensor, coords_h: Tensor, coords_w: Tensor, coords_d: Tensor,
                     position_range) -> Tensor:
    """petr_head.py:303-325.  Returns normalised coords3d, shape (B,N,W,H,D,3) fp32.

    The 4x4 . 4x1 product is evaluated as ((m0*p0 + m1*p1) + m2*p2) + m3*p3 with separately
    rounded fp32 multiplies and adds, which is bit-identical to torch.matmul on CPU for this
    shape (verified against the reference in tests/test_oracle_vs_reference.py)."""
    eps = 1e-5
    W, H, D = coords_w.numel(), coords_h.numel(), coords_d.numel()
    u = coords_w.view(W, 1, 1).expand(W, H, D)
    v = coords_h.view(1, H, 1).expand(W, H, D)
    d = coords_d.view(1, 1, D).expand(W, H, D)
    dm = torch.maximum(d, torch.full_like(d, eps))
    p = [u * dm, v * dm, d, torch.ones_like(d)]                     # (W,H,D) each
    M = img2lidars[:, :, None, None, None]                          # (B,N,1,1,1,4,4)
    rows = []
    for i in range(3):
        acc = M[..., i, 0] * p[0]
        acc = acc + M[..., i, 1] * p[1]
        acc = acc + M[..., i, 2] * p[2]
        acc = acc + M[..., i, 3] * p[3]
        rows.append(acc)
    c = torch.stack(rows, -1)                                       # (B,N,W,H,D,3)
    pr = position_range
    c[..., 0:1] = (c[..., 0:1] - pr[0]) / (pr[3] - pr[0])
    c[..., 1:2] = (c[..., 1:2] - pr[1]) / (pr[4] - pr[1])
    c[..., 2:3] = (c[..., 2:3] - pr[2]) / (pr[5] - pr[2])
    return c


def position_features(img_metas, B, N, H, W, masks: Optional[Tensor], depth_num=64,
                      depth_start=1.0, position_range=(-61.2, -61.2, -10.0, 61.2, 61.2, 10.0),
                      LID=True):
    """petr_head.py:286-331 up to (excluding) the MLP.

    Returns (feat (B*N, 3*D, H, W) fp32 = inverse_sigmoid(coords), channel = 3*d+c;
             coords_mask (B,N,H,W) bool; raw normalised coords (B,N,W,H,D,3))."""
    pad_h, pad_w, _ = img_metas[0]["pad_shape"][0]
    ch, cw, cd = frustum_axes(H, W, pad_h, pad_w, depth_num, depth_start, position_range, LID)
    D = cd.numel()
    c3 = frustum_coords3d(img2lidar_matrices(img_metas), ch, cw, cd, position_range)
    cmask = (c3 > 1.0) | (c3 < 0.0)
    cmask = cmask.flatten(-2).sum(-1) > (D * 0.5)                    # (B,N,W,H)
    cmask = cmask.permute(0, 1, 3, 2)
    if masks is not None:
        cmask = masks | cmask
    feat = c3.permute(0, 1, 4, 5, 3, 2).contiguous().view(B * N, -1, H, W)
    return inverse_sigmoid(feat), cmask, c3


def conv1x1_mlp(x: Tensor, w1, b1, w2, b2) -> Tensor:
    """Conv2d(1x1) + ReLU + Conv2d(1x1) (petr_head.py:249-267)."""
    return F.conv2d(F.relu(F.conv2d(x, w1, b1)), w2, b2)


def sine_positional_encoding_3d(mask: Tensor, num_feats=128, temperature=10000, normalize=True,
                                scale=2 * math.pi, eps=1e-6, offset=0.0) -> Tensor:
    """positional_encoding.py:58-100.  mask bool (B,N,H,W) -> (B,N,3*num_feats,H,W)."""
    not_mask = 1 - mask.to(torch.int)
    embeds = [not_mask.cumsum(a, dtype=torch.float32) for a in (1, 2, 3)]   # n, y, x
    if normalize:
        sl = [(slice(None), slice(-1, None)), (slice(None), slice(None), slice(-1, None)),
              (slice(None), slice(None), slice(None), slice(-1, None))]
        embeds = [(e + offset) / (e[s] + eps) * scale for e, s in zip(embeds, sl)]
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    B, N, H, W = mask.shape
    outs = []
    for e in embeds:
        p = e[..., None] / dim_t
        # NB: dim=4 on 5-D tensors => channel order is [64 sines | 64 cosines] per axis, not
        # interleaved (positional_encoding.py:88-96); a reference quirk that must be kept.
        outs.append(torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=4).view(B, N, H, W, -1))
    pos_n, pos_y, pos_x = outs
    return torch.cat((pos_n, pos_y, pos_x), dim=4).permute(0, 1, 4, 2, 3)


def sine_positional_encoding_2d(mask: Tensor, num_feats=128, temperature=10000, normalize=True,
                                scale=2 * math.pi, eps=1e-6, offset=0.0) -> Tensor:
    """mmdet 2.24.1 SinePositionalEncoding.forward (mmdet/models/utils/positional_encoding.py; un-vendored,
    restated -- parity of this piece is unpinned).  mask bool (B,H,W) -> (B, 2*num_feats, H, W) = [pos_y | pos_x]."""
    not_mask = 1 - mask.to(torch.int)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    if normalize:
        y_embed = (y_embed + offset) / (y_embed[:, -1:, :] + eps) * scale
        x_embed = (x_embed + offset) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    B, H, W = mask.shape
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def learned_positional_encoding_3d(sd, p: str, mask: Tensor) -> Tensor:
    """positional_encoding.py:137-160.  mask (B,h,w) -> (B, 2F, h, w) = [col_embed(x) | row_embed(y)]."""
    h, w = mask.shape[-2:]
    x_embed = sd[p + "col_embed.weight"][:w]
    y_embed = sd[p + "row_embed.weight"][:h]
    pos = torch.cat((x_embed.unsqueeze(0).repeat(h, 1, 1), y_embed.unsqueeze(1).repeat(1, w, 1)), dim=-1)
    return pos.permute(2, 0, 1).unsqueeze(0).repeat(mask.shape[0], 1, 1, 1)


def cpfpn_forward(sd, inputs: Sequence[Tensor], start_level: int = 0) -> List[Tensor]:
    """CPFPN.forward cp_fpn.py:159-210 (no norm / activation, nearest top-down, 3x3 conv on level 0 only, no extra
    levels).  inputs NCHW; sd keys `lateral_convs.{i}.conv.*`, `fpn_convs.0.conv.*`."""
    n = len([k for k in sd if k.startswith("lateral_convs.") and k.endswith(".conv.weight")])
    lat = [F.conv2d(inputs[i + start_level], sd[f"lateral_convs.{i}.conv.weight"], sd[f"lateral_convs.{i}.conv.bias"])
           for i in range(n)]
    for i in range(n - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    outs = [F.conv2d(lat[0], sd["fpn_convs.0.conv.weight"], sd["fpn_convs.0.conv.bias"], padding=1)
            if (i == 0 and "fpn_convs.0.conv.weight" in sd) else lat[i] for i in range(n)]
    return outs


def pos2posemb3d(pos: Tensor, num_pos_feats=128, temperature=10000) -> Tensor:
    """petr_head.py:31-43.  (.., 3) -> (.., 384) in (y, x, z) order."""
    pos = pos * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    emb = []
    for a in (1, 0, 2):
        p = pos[..., a, None] / dim_t
        emb.append(torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2))
    return torch.cat(emb, dim=-1)


# --------------------------------------------------------------------------------------
# transformer
# --------------------------------------------------------------------------------------
def multihead_attention(sd: Dict[str, Tensor], prefix: str, query, key, value, identity,
                        query_pos, key_pos, key_padding_mask, num_heads=8, attn_mask=None) -> Tensor:
    """petr_transformer.py:280-367 (== mmcv MultiheadAttention), eval mode, seq-first.
    Calls the same torch functional as nn.MultiheadAttention (need_weights=True as the
    reference leaves it)."""
    q = query + query_pos if query_pos is not None else query
    k = key + key_pos if key_pos is not None else key
    E = q.shape[-1]
    out, _ = F.multi_head_attention_forward(
        q, k, value, E, num_heads,
        sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"],
        None, None, False, 0.0,
        sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"],
        training=False, key_padding_mask=key_padding_mask, need_weights=True, attn_mask=attn_mask)
    return identity + out


def decoder_layer(sd, p: str, query, memory, query_pos, key_pos, key_padding_mask,
                  num_heads=8, taps: Optional[dict] = None) -> Tensor:
    """operation_order = (self_attn, norm, cross_attn, norm, ffn, norm); post-norm layer
    (mmcv BaseTransformerLayer; reference near-copy multi_atten_decoder_layer.py:160-293)."""
    E = query.shape[-1]

    def ln(x, i):
        return F.layer_norm(x, (E,), sd[f"{p}norms.{i}.weight"], sd[f"{p}norms.{i}.bias"], 1e-5)

    x = multihead_attention(sd, p + "attentions.0.", query, query, query, query,
                            query_pos, query_pos, None, num_heads)
    if taps is not None:
        taps["self_attn"] = x
    x = ln(x, 0)
    y = multihead_attention(sd, p + "attentions.1.", x, memory, memory, x,
                            query_pos, key_pos, key_padding_mask, num_heads)
    if taps is not None:
        taps["cross_attn"] = y
    y = ln(y, 1)
    h = F.relu(F.linear(y, sd[p + "ffns.0.layers.0.0.weight"], sd[p + "ffns.0.layers.0.0.bias"]))
    z = y + F.linear(h, sd[p + "ffns.0.layers.1.weight"], sd[p + "ffns.0.layers.1.bias"])
    return ln(z, 2)


def transformer(sd, x: Tensor, mask: Tensor, query_embed: Tensor, pos_embed: Tensor,
                num_layers=6, num_heads=8, prefix="transformer.", taps: Optional[list] = None):
    """PETRTransformer.forward petr_transformer.py:69-109 + PETRTransformerDecoder.forward
    :423-447 (return_intermediate=True, post_norm on a copy).  Returns (L,B,Lq,E)."""
    bs, n, c, h, w = x.shape
    memory = x.permute(1, 3, 4, 0, 2).reshape(-1, bs, c)
    pos = pos_embed.permute(1, 3, 4, 0, 2).reshape(-1, bs, c)
    qpos = query_embed.unsqueeze(1).repeat(1, bs, 1)
    kpm = mask.view(bs, -1)
    query = torch.zeros_like(qpos)
    E = c
    inter = []
    for l in range(num_layers):
        t = {} if taps is not None else None
        query = decoder_layer(sd, f"{prefix}decoder.layers.{l}.", query, memory, qpos, pos, kpm,
                              num_heads, t)
        if taps is not None:
            t["out"] = query
            taps.append(t)
        inter.append(F.layer_norm(query, (E,), sd[prefix + "decoder.post_norm.weight"],
                                  sd[prefix + "decoder.post_norm.bias"], 1e-5))
    return torch.stack(inter).transpose(1, 2)


# --------------------------------------------------------------------------------------
# branches
# --------------------------------------------------------------------------------------
def normed_linear(x: Tensor, w: Tensor, b: Optional[Tensor], tempearture=20, power=1.0, eps=1e-6) -> Tensor:
    """mmdet 2.24.1 NormedLinear.forward (mmdet/models/utils/normed_predictor.py; un-vendored, restated -- parity of
    this piece is unpinned): both norms run over dim 1, which for the head's (B, Lq, E) activations is the query axis."""
    w_ = w / (w.norm(dim=1, keepdim=True).pow(power) + eps)
    x_ = x / (x.norm(dim=1, keepdim=True).pow(power) + eps)
    return F.linear(x_ * tempearture, w_, b)


def cls_branch(sd, p: str, x: Tensor, num_reg_fcs=2, normedlinear=False) -> Tensor:
    """petr_head.py:226-236: (Linear, LayerNorm, ReLU) x num_reg_fcs, Linear | NormedLinear."""
    E = x.shape[-1]
    for i in range(num_reg_fcs):
        x = F.linear(x, sd[f"{p}{3 * i}.weight"], sd[f"{p}{3 * i}.bias"])
        x = F.relu(F.layer_norm(x, (E,), sd[f"{p}{3 * i + 1}.weight"], sd[f"{p}{3 * i + 1}.bias"], 1e-5))
    j = 3 * num_reg_fcs
    if normedlinear:
        return normed_linear(x, sd[f"{p}{j}.weight"], sd[f"{p}{j}.bias"])
    return F.linear(x, sd[f"{p}{j}.weight"], sd[f"{p}{j}.bias"])


def reg_branch(sd, p: str, x: Tensor, num_reg_fcs=2) -> Tensor:
    """petr_head.py:238-243: (Linear, ReLU) x num_reg_fcs, Linear."""
    for i in range(num_reg_fcs):
        x = F.relu(F.linear(x, sd[f"{p}{2 * i}.weight"], sd[f"{p}{2 * i}.bias"]))
    j = 2 * num_reg_fcs
    return F.linear(x, sd[f"{p}{j}.weight"], sd[f"{p}{j}.bias"])


def reg_layer_multi(sd, p: str, x: Tensor, num_reg_fcs=2, num_tasks=5) -> Tensor:
    """PETRv2 RegLayer petrv2_head.py:63-95: shared trunk + per-task heads, concatenated."""
    for i in range(num_reg_fcs):
        x = F.relu(F.linear(x, sd[f"{p}reg_branch.{3 * i}.weight"], sd[f"{p}reg_branch.{3 * i}.bias"]))
    outs = []
    for t in range(num_tasks):
        h = F.relu(F.linear(x, sd[f"{p}task_heads.{t}.0.weight"], sd[f"{p}task_heads.{t}.0.bias"]))
        outs.append(F.linear(h, sd[f"{p}task_heads.{t}.2.weight"], sd[f"{p}task_heads.{t}.2.bias"]))
    return torch.cat(outs, -1)


# --------------------------------------------------------------------------------------
# whole head
# --------------------------------------------------------------------------------------
DEFAULT_CFG = dict(
    num_query=900, depth_num=64, depth_start=1.0, LID=True, with_position=True,
    with_multiview=True, position_range=(-61.2, -61.2, -10.0, 61.2, 61.2, 10.0),
    pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), num_layers=6, num_heads=8,
    num_reg_fcs=2, with_fpe=False, with_time=False, with_multi=False,
    num_feats=128, normalize=True)


def head_forward(sd: Dict[str, Tensor], feats: Tensor, img_metas: List[dict],
                 cfg: Optional[dict] = None, taps: Optional[dict] = None):
    """PETRHead.forward petr_head.py:366-468 / PETRv2Head.forward petrv2_head.py:429-540,
    eval mode, fp32, CPU.  ``feats`` is mlvl_feats[position_level], (B,N,Cin,H,W).
    Returns dict(all_cls_scores (L,B,Lq,C), all_bbox_preds (L,B,Lq,10))."""
    c = dict(DEFAULT_CFG)
    if cfg:
        c.update(cfg)
    B, N, Cin, H, W = feats.shape
    masks = key_padding_masks(img_metas, B, N, H, W)
    x = F.conv2d(feats.flatten(0, 1), sd["input_proj.weight"], sd["input_proj.bias"])
    x = x.view(B, N, *x.shape[-3:])
    # positional encoding of the key masks: SinePositionalEncoding3D over all views (with_multiview), else a 2-D
    # encoding evaluated per camera and concatenated (petr_head.py:403-409,:414-420)
    kind = c.get("pos_kind", "sine3d")
    if c["with_multiview"]:
        assert kind == "sine3d"
        sin = sine_positional_encoding_3d(masks, c["num_feats"], normalize=c["normalize"]).flatten(0, 1)
    else:
        per_cam = []
        for i in range(N):
            if kind == "sine2d":
                per_cam.append(sine_positional_encoding_2d(masks[:, i], c["num_feats"], normalize=c["normalize"]).unsqueeze(1))
            else:
                per_cam.append(learned_positional_encoding_3d(sd, "positional_encoding.", masks[:, i]).unsqueeze(1))
        sin = torch.cat(per_cam, 1).flatten(0, 1)
    if c["with_multiview"] or c["with_position"]:
        sin = conv1x1_mlp(sin, sd["adapt_pos3d.0.weight"], sd["adapt_pos3d.0.bias"],
                          sd["adapt_pos3d.2.weight"], sd["adapt_pos3d.2.bias"])
    sin = sin.view(x.shape)
    if taps is not None:
        taps["masks"] = masks
        taps["memory"] = x
    if c["with_position"]:
        feat, cmask, _ = position_features(img_metas, B, N, H, W, masks, c["depth_num"],
                                           c["depth_start"], c["position_range"], c["LID"])
        pos = conv1x1_mlp(feat, sd["position_encoder.0.weight"], sd["position_encoder.0.bias"],
                          sd["position_encoder.2.weight"], sd["position_encoder.2.bias"])
        if taps is not None:
            taps["coords_position_embeding"] = pos.view(x.shape)
            taps["coords_mask"] = cmask
        if c["with_fpe"]:
            # SELayer petrv2_head.py:48-60, called at petrv2_head.py:464-466
            xf = x.flatten(0, 1)
            se = F.conv2d(F.relu(F.conv2d(xf, sd["fpe.conv_reduce.weight"], sd["fpe.conv_reduce.bias"])),
                          sd["fpe.conv_expand.weight"], sd["fpe.conv_expand.bias"])
            pos = pos * torch.sigmoid(se)
        pos_embed = pos.view(x.shape) + sin
    else:
        pos_embed = sin
    ref = sd["reference_points.weight"]
    qe = pos2posemb3d(ref)
    qe = F.linear(F.relu(F.linear(qe, sd["query_embedding.0.weight"], sd["query_embedding.0.bias"])),
                  sd["query_embedding.2.weight"], sd["query_embedding.2.bias"])
    if taps is not None:
        taps["pos_embed"] = pos_embed
        taps["query_embeds"] = qe
        taps["layers"] = []
    outs_dec = transformer(sd, x, masks, qe, pos_embed, c["num_layers"], c["num_heads"],
                           taps=taps["layers"] if taps is not None else None)
    outs_dec = torch.nan_to_num(outs_dec)
    if taps is not None:
        taps["outs_dec"] = outs_dec
    reference = inverse_sigmoid(ref.unsqueeze(0).repeat(B, 1, 1))
    mean_ts = None
    if c["with_time"]:
        ts = torch.tensor(np.asarray([np.asarray(m["timestamp"]) for m in img_metas]),
                          dtype=torch.float32).view(B, -1, 6)
        mean_ts = (ts[:, 1, :] - ts[:, 0, :]).mean(-1)
    cls_out, reg_out = [], []
    for l in range(outs_dec.shape[0]):
        cls_out.append(cls_branch(sd, f"cls_branches.{l}.", outs_dec[l], c["num_reg_fcs"],
                                  normedlinear=c.get("normedlinear", False)))
        if c["with_multi"]:
            tmp = reg_layer_multi(sd, f"reg_branches.{l}.", outs_dec[l], c["num_reg_fcs"])
        else:
            tmp = reg_branch(sd, f"reg_branches.{l}.", outs_dec[l], c["num_reg_fcs"])
        tmp[..., 0:2] = (tmp[..., 0:2] + reference[..., 0:2]).sigmoid()
        tmp[..., 4:5] = (tmp[..., 4:5] + reference[..., 2:3]).sigmoid()
        if mean_ts is not None:
            tmp[..., 8:] = tmp[..., 8:] / mean_ts            # petrv2_head.py:520-521 (B == 1)
        reg_out.append(tmp)
    cls = torch.stack(cls_out)
    box = torch.stack(reg_out)
    pc = c["pc_range"]
    box[..., 0:1] = box[..., 0:1] * (pc[3] - pc[0]) + pc[0]
    box[..., 1:2] = box[..., 1:2] * (pc[4] - pc[1]) + pc[1]
    box[..., 4:5] = box[..., 4:5] * (pc[5] - pc[2]) + pc[2]
    return dict(all_cls_scores=cls, all_bbox_preds=box, enc_cls_scores=None, enc_bbox_preds=None)


# ------------------------------------------------------------------------------------------
# NMSFreeCoder.decode / PETRHead.get_bboxes (post-processing right behind the head)
# ------------------------------------------------------------------------------------------
def denormalize_bbox(n: Tensor) -> Tensor:
    """core/bbox/util.py:63-87: (cx, cy, log w, log l, cz, log h, sin, cos[, vx, vy]) ->
    (cx, cy, cz, w, l, h, rot[, vx, vy])."""
    rot = torch.atan2(n[..., 6:7], n[..., 7:8])
    parts = [n[..., 0:1], n[..., 1:2], n[..., 4:5], n[..., 2:3].exp(), n[..., 3:4].exp(), n[..., 5:6].exp(), rot]
    if n.size(-1) > 8:
        parts += [n[..., 8:9], n[..., 9:10]]
    return torch.cat(parts, dim=-1)


def nms_free_decode(cls_scores: Tensor, bbox_preds: Tensor, max_num: int, num_classes: int,
                    post_center_range, score_threshold=None, bottom_center: bool = False) -> List[dict]:
    """nms_free_coder.py:48-119 (decode_single per sample) [+ petr_head.py:744 when bottom_center].
    cls_scores (B, Q, C) logits, bbox_preds (B, Q, code).  Also returns the query index of every kept box."""
    out = []
    rng = torch.tensor(list(post_center_range), dtype=bbox_preds.dtype)
    for b in range(cls_scores.shape[0]):
        scores, indexs = cls_scores[b].sigmoid().view(-1).topk(max_num)
        labels = indexs % num_classes
        bbox_index = indexs // num_classes
        boxes = denormalize_bbox(bbox_preds[b][bbox_index])
        mask = (boxes[..., :3] >= rng[:3]).all(1) & (boxes[..., :3] <= rng[3:]).all(1)
        if score_threshold:
            mask &= scores > score_threshold
        boxes = boxes[mask]
        if bottom_center:
            boxes[:, 2] = boxes[:, 2] - boxes[:, 5] * 0.5
        out.append(dict(bboxes=boxes, scores=scores[mask], labels=labels[mask], index=bbox_index[mask]))
    return out


# ------------------------------------------------------------------------------------------
# training loss (petr_head.py:470-728) with HungarianAssigner3D
# (core/bbox/assigners/hungarian_assigner_3d.py:60-143).  Differentiable torch restatement (autograd gives the
# reference gradients).  mmdet's FocalLoss / FocalLossCost / L1Loss are un-vendored (mmdet 2.24.1) and restated
# from their published formulas -- that boundary is unpinned; normalize_bbox, BBox3DL1Cost, the assignment and
# loss_single bookkeeping follow the reference's own files.
# ------------------------------------------------------------------------------------------
def normalize_bbox(b: Tensor) -> Tensor:
    """core/bbox/util.py:38-61."""
    cx, cy, cz = b[..., 0:1], b[..., 1:2], b[..., 2:3]
    w, l, h = b[..., 3:4].log(), b[..., 4:5].log(), b[..., 5:6].log()
    rot = b[..., 6:7]
    parts = [cx, cy, w, l, cz, h, rot.sin(), rot.cos()]
    if b.size(-1) > 7:
        parts += [b[..., 7:8], b[..., 8:9]]
    return torch.cat(parts, dim=-1)


def focal_loss_cost(cls_pred: Tensor, gt_labels: Tensor, weight=2.0, alpha=0.25, gamma=2.0, eps=1e-12) -> Tensor:
    """mmdet FocalLossCost._focal_loss_cost."""
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    return (pos[:, gt_labels] - neg[:, gt_labels]) * weight


def hungarian_cost(cls_pred, bbox_pred, gt_bboxes, gt_labels, w_cls=2.0, w_reg=0.25) -> Tensor:
    """hungarian_assigner_3d.py:117-131 (cost matrix as handed to scipy)."""
    cost = focal_loss_cost(cls_pred, gt_labels, w_cls) + \
        torch.cdist(bbox_pred[:, :8], normalize_bbox(gt_bboxes)[:, :8], p=1) * w_reg
    return torch.nan_to_num(cost.detach().cpu(), nan=100.0, posinf=100.0, neginf=-100.0)


def sigmoid_focal_loss(pred: Tensor, target_labels: Tensor, num_classes: int, gamma=2.0, alpha=0.25) -> Tensor:
    """mmdet py_sigmoid_focal_loss (== mmcv's sigmoid_focal_loss op), element-wise; labels == num_classes: background."""
    t = F.one_hot(target_labels, num_classes + 1)[:, :num_classes].type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    return F.binary_cross_entropy_with_logits(pred, t, reduction="none") * fw


def head_loss(out: dict, gt_bboxes: List[Tensor], gt_labels: List[Tensor], num_classes=10, code_weights=None,
              w_cls=2.0, w_bbox=0.25, cost_cls=2.0, cost_reg=0.25, bg_cls_weight=0.0, matches=None):
    """PETRHead.loss petr_head.py:647-728 (loss_single :578-645, get_targets/_get_target_single :470-576).
    gt_bboxes: (G, 9) gravity-centre boxes per image.  Returns (dict of losses, matches[l][b] = (rows, cols))."""
    from scipy.optimize import linear_sum_assignment
    cls_all, box_all = out["all_cls_scores"], out["all_bbox_preds"]
    L, B, Q, C = cls_all.shape
    cw = torch.tensor(code_weights if code_weights is not None else [1.0] * 8 + [0.2, 0.2])
    losses, all_matches = {}, []
    for l in range(L):
        labels = torch.full((B * Q,), num_classes, dtype=torch.long)
        tgt = torch.zeros((B * Q, gt_bboxes[0].shape[1]))
        wts = torch.zeros((B * Q, box_all.shape[-1]))
        npos, ml = 0, []
        for b in range(B):
            if gt_bboxes[b].shape[0] == 0:
                ml.append((np.zeros(0, np.int64), np.zeros(0, np.int64)))
                continue
            if matches is None:
                r, c = linear_sum_assignment(hungarian_cost(cls_all[l, b], box_all[l, b], gt_bboxes[b], gt_labels[b],
                                                            cost_cls, cost_reg))
            else:
                r, c = matches[l][b]
            ml.append((np.asarray(r), np.asarray(c)))
            r, c = torch.as_tensor(r), torch.as_tensor(c)
            labels[b * Q + r] = gt_labels[b][c]
            tgt[b * Q + r] = gt_bboxes[b][c]
            wts[b * Q + r] = 1.0
            npos += len(r)
        all_matches.append(ml)
        cls_avg = max(npos * 1.0 + (B * Q - npos) * bg_cls_weight, 1)
        lc = sigmoid_focal_loss(cls_all[l].reshape(-1, C), labels, num_classes).sum() / cls_avg * w_cls
        ntgt = normalize_bbox(tgt)
        ok = torch.isfinite(ntgt).all(dim=-1)
        wts = wts * cw
        lb = ((box_all[l].reshape(B * Q, -1)[ok, :10] - ntgt[ok, :10]).abs() * wts[ok, :10]).sum() / max(npos, 1) * w_bbox
        pre = "" if l == L - 1 else f"d{l}."
        losses[pre + "loss_cls"], losses[pre + "loss_bbox"] = torch.nan_to_num(lc), torch.nan_to_num(lb)
    # petr_head.py:713-725: the last layer's losses come first in the dict
    order = ["loss_cls", "loss_bbox"] + [f"d{l}.{n}" for l in range(L - 1) for n in ("loss_cls", "loss_bbox")]
    return {k: losses[k] for k in order}, all_matches
