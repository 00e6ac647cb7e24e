"""CPU study (test infrastructure): which bf16 roundings of the tcgen05 path cost how much output
error.  Emulates the roundings inside the oracle's arithmetic.  python -m tests.precision_study"""
import sys
import torch
import torch.nn.functional as F
import numpy as np
import petr_b200 as pb
from oracle import petr_oracle as po
from tests import cases


def bf(x):
    return x.to(torch.bfloat16).float()


def split2(x):
    hi = bf(x)
    return hi, bf(x - hi)


def tf32(x):
    # round-to-nearest to a 10-bit mantissa
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def qlin(x, w, b, mode):
    """query-side linear in the given arithmetic"""
    if mode == "fp32":
        return x @ w.T + b
    if mode == "bf16":
        return bf(x) @ bf(w).T + b
    if mode == "tf32":
        return tf32(x) @ tf32(w).T + b
    if mode == "bf16x3":
        xh, xl = split2(x); wh, wl = split2(w)
        return xh @ wh.T + xl @ wh.T + xh @ wl.T + b
    raise ValueError(mode)


def forward(sd, feats, metas, cfg, fl):
    c = dict(po.DEFAULT_CFG); c.update(cfg)
    B, N, Cin, H, W = feats.shape
    masks = po.key_padding_masks(metas, B, N, H, W)
    x = F.conv2d(feats.flatten(0, 1), sd["input_proj.weight"], sd["input_proj.bias"])
    mem = x.permute(0, 2, 3, 1).reshape(-1, 256)                       # tokens
    feat, _, _ = po.position_features(metas, B, N, H, W, masks, 64, 1.0, c["position_range"], True)
    f = feat.permute(0, 2, 3, 1).reshape(-1, 192)
    w1 = sd["position_encoder.0.weight"].flatten(1); b1 = sd["position_encoder.0.bias"]
    w2 = sd["position_encoder.2.weight"].flatten(1); b2 = sd["position_encoder.2.bias"]

    def mlp(f, w1, b1, w2, b2, mode):
        if mode == "fp32":
            return F.relu(f @ w1.T + b1) @ w2.T + b2
        if mode == "bf16":
            h = bf(F.relu(bf(f) @ bf(w1).T + b1))
            return h @ bf(w2).T + b2
        if mode == "split_in":      # input split hi+lo, hidden bf16
            fh, fl_ = split2(f)
            h = bf(F.relu(fh @ bf(w1).T + fl_ @ bf(w1).T + b1))
            return h @ bf(w2).T + b2
        if mode == "split_all":     # inputs and hidden split (weights bf16)
            fh, fl_ = split2(f)
            h = F.relu(fh @ bf(w1).T + fl_ @ bf(w1).T + b1)
            hh, hl = split2(h)
            return hh @ bf(w2).T + hl @ bf(w2).T + b2
        if mode == "x3":            # bf16x3: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
            fh, fl_ = split2(f); w1h, w1l = split2(w1); w2h, w2l = split2(w2)
            h = F.relu(fh @ w1h.T + fl_ @ w1h.T + fh @ w1l.T + b1)
            hh, hl = split2(h)
            return hh @ w2h.T + hl @ w2h.T + hh @ w2l.T + b2
    pos = mlp(f, w1, b1, w2, b2, fl["pe"])
    sin = po.sine_positional_encoding_3d(masks).permute(0, 1, 3, 4, 2).reshape(-1, 384)
    pos = pos + mlp(sin, sd["adapt_pos3d.0.weight"].flatten(1), sd["adapt_pos3d.0.bias"],
                    sd["adapt_pos3d.2.weight"].flatten(1), sd["adapt_pos3d.2.bias"], fl["ad"])
    ref = sd["reference_points.weight"]
    qe = po.pos2posemb3d(ref)
    qe = F.linear(F.relu(F.linear(qe, sd["query_embedding.0.weight"], sd["query_embedding.0.bias"])),
                  sd["query_embedding.2.weight"], sd["query_embedding.2.bias"])
    Lq = ref.shape[0]; Lk = N * H * W
    kin = (mem + pos).view(B, Lk, 256); vin = mem.view(B, Lk, 256)
    kmask = masks.view(B, Lk)
    X = torch.zeros(B, Lq, 256)
    outs = []
    for l in range(6):
        p = f"transformer.decoder.layers.{l}."
        def ln(x, i): return F.layer_norm(x, (256,), sd[f"{p}norms.{i}.weight"], sd[f"{p}norms.{i}.bias"], 1e-5)
        qm = fl.get("qside", "fp32")
        ws = sd[p + "attentions.0.attn.in_proj_weight"]; bs = sd[p + "attentions.0.attn.in_proj_bias"]
        sq = qlin(X + qe, ws[:512], bs[:512], qm); sv = qlin(X, ws[512:], bs[512:], qm)
        sqh = (sq[..., :256] / np.sqrt(32.0)).view(B, Lq, 8, 32).transpose(1, 2)
        skh = sq[..., 256:].view(B, Lq, 8, 32).transpose(1, 2); svh = sv.view(B, Lq, 8, 32).transpose(1, 2)
        sam = fl.get("sattn", "fp32")
        if sam == "tf32qk_bf16pv":
            Ss = tf32(sqh) @ tf32(skh).transpose(-1, -2)
            mm = Ss.max(-1, keepdim=True).values; pu = torch.exp(Ss - mm)
            so = (bf(pu) @ bf(svh)) / pu.sum(-1, keepdim=True)
        else:
            so = torch.softmax(sqh @ skh.transpose(-1, -2), -1) @ svh
        so = so.transpose(1, 2).reshape(B, Lq, 256)
        X = ln(X + qlin(so, sd[p + "attentions.0.attn.out_proj.weight"], sd[p + "attentions.0.attn.out_proj.bias"], qm), 0)
        w = sd[p + "attentions.1.attn.in_proj_weight"]; b = sd[p + "attentions.1.attn.in_proj_bias"]
        q = qlin(X + qe, w[:256], b[:256], qm)
        if fl["kv"] == "bf16":
            K = bf(bf(kin) @ bf(w[256:512]).T + b[256:512]); V = bf(bf(vin) @ bf(w[512:]).T + b[512:])
        elif fl["kv"] == "split_in":
            kh, kl = split2(kin); vh, vl = split2(vin)
            K = bf(kh @ bf(w[256:512]).T + kl @ bf(w[256:512]).T + b[256:512])
            V = bf(vh @ bf(w[512:]).T + vl @ bf(w[512:]).T + b[512:])
        elif fl["kv"] == "x3_bf16out":
            kh, kl = split2(kin); vh, vl = split2(vin); wkh, wkl = split2(w[256:512]); wvh, wvl = split2(w[512:])
            K = bf(kh @ wkh.T + kl @ wkh.T + kh @ wkl.T + b[256:512])
            V = bf(vh @ wvh.T + vl @ wvh.T + vh @ wvl.T + b[512:])
        elif fl["kv"] == "kx3_vbf16":       # K: bf16x3 projection of fp32 kin, kept fp32; V: bf16 projection, bf16
            K = kin @ w[256:512].T + b[256:512]; V = bf(bf(vin) @ bf(w[512:]).T + b[512:])
        elif fl["kv"] == "kx3tf32_vbf16":
            K = tf32(kin @ w[256:512].T + b[256:512]); V = bf(bf(vin) @ bf(w[512:]).T + b[512:])
        elif fl["kv"] == "kx3bf16_vbf16":   # K x3 projection but rounded to bf16 at the end
            K = bf(kin @ w[256:512].T + b[256:512]); V = bf(bf(vin) @ bf(w[512:]).T + b[512:])
        elif fl["kv"] == "kbf16in_f32out":  # K: bf16 projection inputs, fp32 output
            K = bf(kin) @ bf(w[256:512]).T + b[256:512]; V = bf(bf(vin) @ bf(w[512:]).T + b[512:])
        elif fl["kv"] == "tf32k_bf16v":     # bf16 projection GEMM, K kept fp32 (tf32-rounded), V bf16
            K = tf32(bf(kin) @ bf(w[256:512]).T + b[256:512]); V = bf(bf(vin) @ bf(w[512:]).T + b[512:])
        else:
            K = kin @ w[256:512].T + b[256:512]; V = vin @ w[512:].T + b[512:]
        qs = q / np.sqrt(32.0)
        if fl["qk"] == "bf16":
            qs = bf(qs)
        elif fl["qk"] == "tf32":
            qs = tf32(qs)
        qh = qs.view(B, Lq, 8, 32).transpose(1, 2); Kh = K.view(B, Lk, 8, 32).transpose(1, 2)
        Vh = V.view(B, Lk, 8, 32).transpose(1, 2)
        S = qh @ Kh.transpose(-1, -2)
        S = S.masked_fill(kmask[:, None, None, :], float("-inf"))
        Pm = torch.softmax(S, -1)
        if fl["p"] == "bf16":
            m = S.max(-1, keepdim=True).values
            pu = torch.exp(S - m); l_ = pu.sum(-1, keepdim=True)
            O = (bf(pu) @ Vh) / l_
        else:
            O = Pm @ Vh
        O = O.transpose(1, 2).reshape(B, Lq, 256)
        X = ln(X + qlin(O, sd[p + "attentions.1.attn.out_proj.weight"], sd[p + "attentions.1.attn.out_proj.bias"], qm), 1)
        h = F.relu(qlin(X, sd[p + "ffns.0.layers.0.0.weight"], sd[p + "ffns.0.layers.0.0.bias"], qm))
        X = ln(X + qlin(h, sd[p + "ffns.0.layers.1.weight"], sd[p + "ffns.0.layers.1.bias"], qm), 2)
        outs.append(F.layer_norm(X, (256,), sd["transformer.decoder.post_norm.weight"], sd["transformer.decoder.post_norm.bias"], 1e-5))
    outs = torch.nan_to_num(torch.stack(outs))
    reference = po.inverse_sigmoid(ref.unsqueeze(0).repeat(B, 1, 1))
    cls_o, reg_o = [], []
    for l in range(6):
        cls_o.append(po.cls_branch(sd, f"cls_branches.{l}.", outs[l]))
        t = po.reg_branch(sd, f"reg_branches.{l}.", outs[l])
        t[..., 0:2] = (t[..., 0:2] + reference[..., 0:2]).sigmoid(); t[..., 4:5] = (t[..., 4:5] + reference[..., 2:3]).sigmoid()
        reg_o.append(t)
    cls = torch.stack(cls_o); box = torch.stack(reg_o)
    box[..., 0] = box[..., 0] * 102.4 - 51.2; box[..., 1] = box[..., 1] * 102.4 - 51.2; box[..., 4] = box[..., 4] * 8 - 5
    return cls, box


if __name__ == "__main__":
    names = sys.argv[1:] or ["tiny_padded", "tiny"]
    base = dict(pe="fp32", ad="fp32", kv="fp32", qk="fp32", p="fp32")
    variants = {
        "fp32 (emulation self-check)": {},
        "pe bf16": dict(pe="bf16"), "ad bf16": dict(ad="bf16"), "kv bf16": dict(kv="bf16"),
        "q bf16": dict(qk="bf16"), "p bf16": dict(p="bf16"),
        "ALL bf16 (current)": dict(pe="bf16", ad="bf16", kv="bf16", qk="bf16", p="bf16"),
        "pe/ad split_in, rest bf16": dict(pe="split_in", ad="split_in", kv="bf16", qk="bf16", p="bf16"),
        "pe/ad split_all, rest bf16": dict(pe="split_all", ad="split_all", kv="bf16", qk="bf16", p="bf16"),
        "pe/ad x3, rest bf16": dict(pe="x3", ad="x3", kv="bf16", qk="bf16", p="bf16"),
        "pe/ad x3, kv split_in": dict(pe="x3", ad="x3", kv="split_in", qk="bf16", p="bf16"),
        "pe/ad x3, kv x3": dict(pe="x3", ad="x3", kv="x3_bf16out", qk="bf16", p="bf16"),
        "pe/ad fp32, kv x3": dict(kv="x3_bf16out", qk="bf16", p="bf16"),
    }
    for name in names:
        case = cases.CASES[name]
        wl, feats, metas = cases.inputs(case)
        head = pb.build_head(cases.head_cfg(case))
        sd = cases.weights(case, head)
        with torch.no_grad():
            ref = po.head_forward(sd, feats, metas, cases.oracle_cfg(case))
            for vn, v in variants.items():
                fl = dict(base); fl.update(v)
                cls, box = forward(sd, feats, metas, cases.oracle_cfg(case), fl)
                ec = (cls - ref["all_cls_scores"]).abs().max().item()
                eb = (box - ref["all_bbox_preds"]).abs()
                print(f"{name:12s} {vn:32s} cls {ec:.4f}  box {eb.max().item():.4f}  cx/cy {eb[...,0:2].max().item():.4f} others {eb[...,2:].max().item():.4f}", flush=True)
