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
