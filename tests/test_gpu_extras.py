"""GPU parity of the pieces either side of the main path: the reference branches no shipped config takes
(with_multiview=False with a 2-D positional encoding, with_position=False, normedlinear=True, attn_mask), the
stand-alone attention / FFN / transformer bricks on the tcgen05 path, the PETRv2 SE gate as a GEMM epilogue, the CPFPN
neck tail and DepthrHead.position_embeding.  References: the CPU oracle, itself pinned
against the unmodified reference in tests/test_oracle_vs_reference.py."""
import numpy as np
import pytest
import torch

import petr_b200 as pb
from oracle import petr_oracle as po
from petr_b200 import ops, synth
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

LEARNED = dict(type='LearnedPositionalEncoding3D', num_feats=128, row_num_embed=50, col_num_embed=50)
SINE2D = dict(type='SinePositionalEncoding', num_feats=128, normalize=True)

VARIANTS = {
    # name: (head cfg overrides, oracle cfg overrides)
    "learned_pos": (dict(with_multiview=False, positional_encoding=LEARNED), dict(with_multiview=False, pos_kind="learned")),
    "learned_only": (dict(with_multiview=False, with_position=False, positional_encoding=LEARNED),
                     dict(with_multiview=False, with_position=False, pos_kind="learned")),
    "sine2d_pos": (dict(with_multiview=False, positional_encoding=SINE2D), dict(with_multiview=False, pos_kind="sine2d")),
    "sine3d_only": (dict(with_position=False), dict(with_position=False)),
    "normedlinear": (dict(normedlinear=True), dict(normedlinear=True)),
}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("case_name", ["small", "small_padded"])
def test_head_config_branches(variant, case_name):
    """petr_head.py:396-420 branches and NormedLinear against the oracle: fp32 path <= 2e-3, bf16 path <= 2e-2."""
    case = cases.CASES[case_name]
    hc, oc = VARIANTS[variant]
    wl, feats, metas = cases.inputs(case)
    cfg = cases.head_cfg(case)
    cfg.update(hc)
    head = pb.build_head(cfg)
    sd = synth.head_state_dict(head, case["seed"])
    if variant == "normedlinear":      # temperature 20 on unit-norm operands: keep the logits in a sane range
        for k in sd:
            if k.startswith("cls_branches.") and k.endswith(".6.weight"):
                sd[k] = sd[k] * 0.05
    head.load_state_dict(sd, strict=True)
    head = head.eval().to(DEV)
    ocfg = cases.oracle_cfg(case)
    ocfg.update(oc)
    with torch.no_grad():
        ref = po.head_forward(sd, feats, metas, ocfg)
    x = feats.to(DEV)
    for prec, tol in (("fp32", 2e-3), ("bf16", 2e-2)):
        head.precision = prec
        out = head([x], metas)
        for k in ("all_cls_scores", "all_bbox_preds"):
            err = (out[k].cpu() - ref[k]).abs().max().item()
            if variant == "normedlinear" and k == "all_cls_scores":
                # temperature 20 on unit-norm operands: logits of magnitude ~20; the bound is relative to their scale.
                # NormedLinear divides every channel by its norm over the QUERY axis (x.norm(dim=1) on a 3-D input): a
                # channel that ReLU leaves (almost) dead for all queries is divided by ~0, so bf16-level differences
                # upstream are amplified without bound -- the op is ill-conditioned, the bf16 bound is not applicable
                if prec == "bf16":
                    continue
                err /= max(1.0, ref[k].abs().max().item())
            print(f"{variant} {case_name} {prec} {k}: {err:.2e}")
            assert np.isfinite(err) and err <= tol, (variant, prec, k, err)


def test_positional_encoding_modules_2d():
    g = torch.Generator().manual_seed(0)
    mask = torch.rand(3, 12, 20, generator=g) < 0.2
    mask[1, 8:] = True
    mask[2, :, 15:] = True
    pe = pb.SinePositionalEncoding(num_feats=128, normalize=True).to(DEV)
    got = pe(mask.to(DEV))
    ref = po.sine_positional_encoding_2d(mask, 128, normalize=True)
    assert got.shape == ref.shape and (got.cpu() - ref).abs().max() <= 2e-6
    pe = pb.SinePositionalEncoding(num_feats=64, normalize=False, temperature=20).to(DEV)
    assert (pe(mask.to(DEV)).cpu() - po.sine_positional_encoding_2d(mask, 64, 20, normalize=False)).abs().max() <= 2e-5
    le = pb.LearnedPositionalEncoding3D(num_feats=128)
    torch.nn.init.uniform_(le.row_embed.weight, generator=g)
    torch.nn.init.uniform_(le.col_embed.weight, generator=g)
    le = le.to(DEV)
    sd = {"pe." + k: v.cpu() for k, v in le.state_dict().items()}
    assert torch.equal(le(mask.to(DEV)).cpu(), po.learned_positional_encoding_3d(sd, "pe.", mask))
    with pytest.raises(RuntimeError):
        le(torch.zeros(1, 60, 20, dtype=torch.bool, device=DEV))       # 60 rows > row_num_embed


def test_attn_mask_and_tc_bricks():
    """Stand-alone PETRMultiheadAttention / FFN / PETRTransformer: attn_mask (bool, additive, per-head) on the fp32
    kernels == torch; default 'bf16' precision runs the tcgen05 kernels within bf16 tolerance."""
    case = cases.CASES["small_padded"]
    wl, feats, metas = cases.inputs(case)
    head = pb.build_head(cases.head_cfg(case))
    head.load_state_dict(cases.weights(case, head), strict=True)
    head = head.eval().to(DEV)
    sd = {k: v.cpu() for k, v in head.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    B, Lq, Lk, E = 2, 70, 300, 256
    query, qpos = torch.randn(Lq, B, E, generator=g), torch.randn(Lq, B, E, generator=g)
    key, kpos = torch.randn(Lk, B, E, generator=g), torch.randn(Lk, B, E, generator=g)
    kpm = torch.rand(B, Lk, generator=g) < 0.3
    att = head.transformer.decoder.layers[1].attentions[1]
    p = "transformer.decoder.layers.1.attentions.1."
    masks = [torch.rand(Lq, Lk, generator=g) < 0.4, torch.randn(Lq, Lk, generator=g),
             torch.randn(B * 8, Lq, Lk, generator=g)]
    att.precision = "fp32"
    for am in masks:
        ref = po.multihead_attention(sd, p, query, key, key, query, qpos, kpos, kpm, attn_mask=am)
        got = att(query.to(DEV), key.to(DEV), None, None, query_pos=qpos.to(DEV), key_pos=kpos.to(DEV),
                  attn_mask=am.to(DEV), key_padding_mask=kpm.to(DEV))
        assert (got.cpu() - ref).abs().max() <= 2e-4
    # tensor-core path (default precision)
    att.precision = "bf16"
    ops.reset_launch_count()
    ref = po.multihead_attention(sd, p, query, key, key, query, qpos, kpos, kpm)
    got = att(query.to(DEV), key.to(DEV), None, None, query_pos=qpos.to(DEV), key_pos=kpos.to(DEV),
              key_padding_mask=kpm.to(DEV))
    assert (got.cpu() - ref).abs().max() <= 2e-2
    sa = head.transformer.decoder.layers[0].attentions[0]
    assert sa.precision == "bf16"
    ref = po.multihead_attention(sd, "transformer.decoder.layers.0.attentions.0.", query, query, query, query, qpos, qpos, None)
    got = sa(query.to(DEV), query_pos=qpos.to(DEV))
    assert (got.cpu() - ref).abs().max() <= 1e-3            # x3 (hi/lo split) kernel: fp32-class
    ffn = head.transformer.decoder.layers[0].ffns[0]
    pf = "transformer.decoder.layers.0.ffns.0.layers."
    x = torch.randn(Lq, B, E, generator=g)
    ref = x + torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(x, sd[pf + "0.0.weight"], sd[pf + "0.0.bias"])),
                                         sd[pf + "1.weight"], sd[pf + "1.bias"])
    assert (ffn(x.to(DEV)).cpu() - ref).abs().max() <= 1e-3
    # whole transformer, batch-first feature maps, on the tensor-core bricks
    Bf, N, _, H, W = feats.shape
    xm = torch.randn(Bf, N, E, H, W, generator=g)
    pos_embed = torch.randn(Bf, N, E, H, W, generator=g)
    mk = po.key_padding_masks(metas, Bf, N, H, W)
    qe = torch.randn(64, E, generator=g)
    ref = torch.nan_to_num(po.transformer(sd, xm, mk, qe, pos_embed))
    out, mem = head.transformer(xm.to(DEV), mk.to(DEV), qe.to(DEV), pos_embed.to(DEV), head.reg_branches)
    ok = ~torch.isnan(out[:, :, 0, 0]).cpu()
    assert (torch.nan_to_num(out.cpu()) - ref)[ok].abs().max() <= 2e-2
    assert ops.launch_count() > 100


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1.5e-2)])
def test_cpfpn(precision, tol):
    """CPFPN (cp_fpn.py:159-210): laterals, nearest top-down add, 3x3 conv on level 0; relative to the output scale."""
    g = torch.Generator().manual_seed(3)
    neck = pb.build_neck(dict(type='CPFPN', in_channels=[128, 192], out_channels=256, num_outs=2))
    neck.init_weights()
    for p_ in neck.parameters():
        if p_.dim() == 1:
            torch.nn.init.normal_(p_, std=0.1, generator=g)
    sd = {k: v.detach().clone() for k, v in neck.state_dict().items()}
    xs = [torch.randn(3, 128, 13, 22, generator=g), torch.randn(3, 192, 7, 11, generator=g)]
    ref = po.cpfpn_forward(sd, xs)
    neck = neck.eval().to(DEV)
    neck.precision = precision
    ops.reset_launch_count()
    outs = neck([x.to(DEV) for x in xs])
    assert ops.launch_count() >= 6
    assert len(outs) == 2
    for o, r in zip(outs, ref):
        assert o.shape == r.shape
        err = (o.cpu() - r).abs().max().item() / r.abs().max().item()
        print(f"cpfpn {precision}: rel err {err:.2e}")
        assert err <= tol, err
    # the PETR configuration's channel counts / strides (vov p4: 768 @ stride 16, 1024 @ stride 32), one sample
    neck = pb.build_neck(dict(type='CPFPN', in_channels=[768, 1024], out_channels=256, num_outs=2))
    neck.init_weights()
    sd = {k: v.detach().clone() for k, v in neck.state_dict().items()}
    xs = [torch.randn(6, 768, 20, 50, generator=g), torch.randn(6, 1024, 10, 25, generator=g)]
    ref = po.cpfpn_forward(sd, xs)
    neck = neck.eval().to(DEV)
    neck.precision = precision
    for o, r in zip(neck([x.to(DEV) for x in xs]), ref):
        assert (o.cpu() - r).abs().max().item() / r.abs().max().item() <= tol


def test_depthr_head_position_embeding():
    """DepthrHead.position_embeding (depthr_head.py:338-386) == PETRHead's arithmetic: built from the reference's
    Depthr constructor kwargs, checked against the oracle (1e-5 of scale, coords_mask bit-exact)."""
    case = cases.CASES["small_padded"]
    wl, feats, metas = cases.inputs(case)
    cfg = cases.head_cfg(case)
    cfg.update(type='DepthrHead', depth_gt_encoder=dict(type='DepthGTEncoder', num_depth_bins=80),
               transformer=dict(type='DepthrTransformer', decoder=dict(type='DepthrTransformerDecoder')))
    head = pb.build_head(cfg)
    assert head.transformer is None and head.transformer_cfg["type"] == 'DepthrTransformer'
    g = torch.Generator().manual_seed(5)
    for p_ in head.position_encoder.parameters():
        torch.nn.init.normal_(p_, std=0.05, generator=g)
    head = head.eval().to(DEV)
    sd = {k: v.cpu() for k, v in head.state_dict().items()}
    B, N, _, H, W = feats.shape
    masks = po.key_padding_masks(metas, B, N, H, W)
    feat, cmask, _ = po.position_features(metas, B, N, H, W, masks, 64, 1.0, tuple(cfg["position_range"]), True)
    ref = po.conv1x1_mlp(feat, sd["position_encoder.0.weight"], sd["position_encoder.0.bias"],
                         sd["position_encoder.2.weight"], sd["position_encoder.2.bias"]).view(B, N, 256, H, W)
    pos, cm = head.position_embeding([feats.to(DEV)], metas, masks.to(DEV))
    assert torch.equal(cm.cpu(), cmask)
    assert (pos.cpu() - ref).abs().max() <= 1e-5 * ref.abs().max()
    with pytest.raises(NotImplementedError):
        head([feats.to(DEV)], metas)


def test_normed_linear_module():
    g = torch.Generator().manual_seed(7)
    nl = pb.NormedLinear(256, 10).to(DEV)
    x3 = torch.randn(2, 50, 256, generator=g)
    x2 = torch.randn(40, 256, generator=g)
    w, b = nl.weight.detach().cpu(), nl.bias.detach().cpu()
    assert (nl(x3.to(DEV)).cpu() - po.normed_linear(x3, w, b)).abs().max() <= 2e-5
    assert (nl(x2.to(DEV)).cpu() - po.normed_linear(x2, w, b)).abs().max() <= 2e-5
