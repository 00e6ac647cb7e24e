"""CPU, build container only: pins oracle/petr_oracle.py against the UNMODIFIED reference
executed through oracle/ref_loader.py.  Skipped where /root/reference is absent (GPU box)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import petr_oracle as po
from oracle import ref_loader as rl
from petr_b200 import synth
from petr_b200.configs import WORKLOADS, petr_head_cfg, petrv2_head_cfg

pytestmark = pytest.mark.skipif(not rl.available(), reason="/root/reference not present")

CFGS = {
    "vov-p4-800x320": "projects/configs/petr/petr_vovnet_gridmask_p4_800x320.py",
    "v2-vov-p4-800x320": "projects/configs/petrv2/petrv2_vovnet_gridmask_p4_800x320.py",
}


@pytest.mark.parametrize("wn", list(CFGS))
def test_head_forward_bit_equal(wn):
    wl = WORKLOADS[wn]
    head = rl.build_reference_head(CFGS[wn], seed=0)
    sd = {k: v.detach() for k, v in head.state_dict().items()}
    metas = synth.make_img_metas(wl, 1, seed=0)
    x = synth.make_features(wl, 1, seed=0)
    v2 = wl.head == "PETRv2Head"
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = head([x], metas)
        masks = po.key_padding_masks(metas, 1, wl.num_cams, wl.feat_h, wl.feat_w)
        rpos, rmask = head.position_embeding([x], metas, masks)
        taps = {}
        out = po.head_forward(sd, x, metas, dict(with_fpe=v2, with_time=v2, with_multi=v2), taps)
    assert torch.equal(rpos, taps["coords_position_embeding"])         # incl. the 4x4 transform order
    assert torch.equal(rmask, taps["coords_mask"])
    assert torch.equal(ref["all_cls_scores"], out["all_cls_scores"])
    assert (ref["all_bbox_preds"] - out["all_bbox_preds"]).abs().max() <= 2e-5


def test_config_builders_match_reference_files():
    pairs = [("petr/petr_vovnet_gridmask_p4_800x320.py", petr_head_cfg(256, with_cp=None)),
             ("petr/petr_vovnet_gridmask_p4_1600x640.py", petr_head_cfg(256)),
             ("petr/petr_r50dcn_gridmask_p4.py", petr_head_cfg(256)),
             ("petr/petr_r50dcn_gridmask_c5.py", petr_head_cfg(2048)),
             ("petrv2/petrv2_vovnet_gridmask_p4_800x320.py", petrv2_head_cfg(256))]
    for f, mine in pairs:
        assert rl.read_head_cfg("projects/configs/" + f) == mine, f


@pytest.mark.parametrize("wn", list(CFGS))
def test_state_dict_contract(wn):
    import petr_b200 as pb
    ref = rl.build_reference_head(CFGS[wn])
    mine = pb.build_head(rl.read_head_cfg(CFGS[wn]))
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert a == b
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(mine.state_dict(), strict=True)


def test_nms_free_coder_and_get_bboxes_match_reference():
    """Next row: the oracle's NMSFreeCoder.decode / PETRHead.get_bboxes restatement against the
    reference's own coder (nms_free_coder.py:48-119, petr_head.py:731-751); the CUDA kernel is then checked
    against the oracle on the GPU box (tests/test_gpu_kernels.py::test_nms_free_decode_kernel)."""
    ns = rl.load_reference()
    g = torch.Generator().manual_seed(3)
    B, Q = 3, 900
    outs = dict(all_cls_scores=torch.randn(6, B, Q, 10, generator=g) - 2.0,
                all_bbox_preds=torch.randn(6, B, Q, 10, generator=g))
    outs["all_bbox_preds"][..., 0:2] *= 40.0
    rng = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    kw = dict(pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], post_center_range=rng,
              max_num=300, voxel_size=[0.2, 0.2, 8], num_classes=10)
    for thr in (None, 0.2):
        ref = ns.NMSFreeCoder(score_threshold=thr, **kw).decode(outs)
        mine = po.nms_free_decode(outs["all_cls_scores"][-1], outs["all_bbox_preds"][-1], 300, 10, rng, thr)
        assert len(ref) == len(mine) == B
        for r, m in zip(ref, mine):
            assert torch.equal(r["labels"], m["labels"]) and torch.equal(r["scores"], m["scores"])
            assert torch.equal(r["bboxes"], m["bboxes"])
    res = po.nms_free_decode(outs["all_cls_scores"][-1], outs["all_bbox_preds"][-1], 300, 10, rng, None, bottom_center=True)
    ref = ns.NMSFreeCoder(**kw).decode(outs)
    for m, r in zip(res, ref):
        exp = r["bboxes"].clone()
        exp[:, 2] = exp[:, 2] - exp[:, 5] * 0.5
        assert torch.equal(m["bboxes"], exp) and torch.equal(m["scores"], r["scores"]) and torch.equal(m["labels"], r["labels"])


@pytest.mark.parametrize("with_position", [True, False])
def test_with_multiview_false_and_normedlinear_branches(with_position):
    """petr_head.py:403-420 (per-camera 2-D encoding; LearnedPositionalEncoding3D is the reference's own 2-D-capable
    module) and normedlinear=True (mmdet NormedLinear stand-in with the published forward)."""
    wl = WORKLOADS["tiny"]
    pe = dict(type='LearnedPositionalEncoding3D', num_feats=128, row_num_embed=50, col_num_embed=50)
    head = rl.build_reference_head(CFGS["vov-p4-800x320"], in_channels=wl.in_channels, num_query=64,
                                   with_multiview=False, with_position=with_position, normedlinear=True,
                                   positional_encoding=pe)
    sd = {k: v.detach() for k, v in head.state_dict().items()}
    metas = synth.make_img_metas(wl, 2, seed=0)
    x = synth.make_features(wl, 2, seed=0)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = head([x], metas)
        out = po.head_forward(sd, x, metas, dict(num_query=64, with_multiview=False, with_position=with_position,
                                                 pos_kind="learned", normedlinear=True))
    assert torch.equal(ref["all_cls_scores"], out["all_cls_scores"])
    assert (ref["all_bbox_preds"] - out["all_bbox_preds"]).abs().max() <= 2e-5
    import petr_b200 as pb
    from petr_b200.configs import petr_head_cfg
    mine = pb.build_head(dict(petr_head_cfg(wl.in_channels), num_query=64, with_multiview=False,
                              with_position=with_position, normedlinear=True, positional_encoding=pe))
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_cpfpn_oracle_bit_equal_and_state_dict():
    ns = rl.load_reference()
    torch.manual_seed(0)
    ref = ns.CPFPN(in_channels=[128, 192], out_channels=256, num_outs=2).eval()
    for p in ref.parameters():
        torch.nn.init.normal_(p, std=0.05)
    xs = [torch.randn(2, 128, 12, 20), torch.randn(2, 192, 6, 10)]
    with torch.no_grad():
        a = ref(xs)
        b = po.cpfpn_forward({k: v.detach() for k, v in ref.state_dict().items()}, xs)
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    import petr_b200 as pb
    mine = pb.build_neck(dict(type='CPFPN', in_channels=[128, 192], out_channels=256, num_outs=2))
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in ref.state_dict().items()}
