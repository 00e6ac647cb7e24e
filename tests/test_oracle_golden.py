"""CPU: the travelling oracle (oracle/petr_oracle.py) must reproduce the committed outputs of
the UNMODIFIED reference (tests/golden/*.npz, made by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import petr_b200 as pb
from oracle import petr_oracle as po
from tests import cases


def _oracle_run(name):
    case = cases.CASES[name]
    wl, feats, metas = cases.inputs(case)
    head = pb.build_head(cases.head_cfg(case))
    sd = cases.weights(case, head)
    taps = {}
    with torch.no_grad():
        out = po.head_forward(sd, feats, metas, cases.oracle_cfg(case), taps)
    return out, taps


@pytest.mark.parametrize("name", [n for n in cases.CASES if n != "vov_800x320"])
def test_oracle_matches_reference_golden(name):
    g = cases.load_golden(name)
    out, taps = _oracle_run(name)
    # same torch build => bit-identical; allow a few ulp for other CPU/MKL code paths
    np.testing.assert_allclose(out["all_cls_scores"].numpy(), g["all_cls_scores"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(out["all_bbox_preds"].numpy(), g["all_bbox_preds"], rtol=0, atol=2e-4)
    assert np.array_equal(taps["coords_mask"].numpy(), g["coords_mask"])
    assert np.array_equal(taps["masks"].numpy(), g["masks"])
    pos = taps["coords_position_embeding"].numpy()
    if "coords_position_embeding" in g:
        assert np.abs(pos - g["coords_position_embeding"]).max() <= 1e-5 * g["pos_scale"]
    else:
        assert np.abs(pos.reshape(-1)[::cases.POS_SAMPLE_STRIDE] - g["pos_sample"]).max() <= 1e-5 * g["pos_scale"]


def test_oracle_matches_reference_golden_800x320():
    g = cases.load_golden("vov_800x320")
    out, taps = _oracle_run("vov_800x320")
    np.testing.assert_allclose(out["all_cls_scores"].numpy(), g["all_cls_scores"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(out["all_bbox_preds"].numpy(), g["all_bbox_preds"], rtol=0, atol=2e-4)
    pos = taps["coords_position_embeding"].numpy().reshape(-1)[::cases.POS_SAMPLE_STRIDE]
    assert np.abs(pos - g["pos_sample"]).max() <= 1e-5 * g["pos_scale"]
    assert np.array_equal(taps["coords_mask"].numpy(), g["coords_mask"])


def test_golden_cases_exercise_the_edge_paths():
    g = cases.load_golden("tiny_padded")
    assert g["masks"].any() and not g["masks"].all()
    assert g["masks"][1].all()                                  # fully masked sample
    assert np.isfinite(g["all_cls_scores"]).all()               # NaN -> 0 path (petr_head.py:435)
    g = cases.load_golden("tiny_range")
    assert g["coords_mask"].any() and not g["coords_mask"].all()


def test_integer_grids_bit_exact():
    """Frustum grids and LID depth bins (bit-exact class) against values captured from the
    reference's own position_embeding."""
    z = cases.load_golden("integer_grids")
    shapes = {"800x320": (20, 50, 320, 800), "c5_1408x512": (16, 44, 512, 1408),
              "1600x640": (40, 100, 640, 1600), "pad928": (58, 100, 928, 1600)}
    for k, (H, W, ph, pw) in shapes.items():
        ch, cw, cd = po.frustum_axes(H, W, ph, pw)
        assert np.array_equal(ch.numpy(), z[f"{k}_coords_h"])
        assert np.array_equal(cw.numpy(), z[f"{k}_coords_w"])
        assert np.array_equal(cd.numpy(), z[f"{k}_coords_d"])
    assert abs(float(z["800x320_coords_d"][63]) - 59.3476906) < 1e-4
    metas = [dict(pad_shape=[(928, 1600, 3)], img_shape=[(900, 1600, 3)])]
    m = po.key_padding_masks(metas, 1, 1, 58, 100)
    assert np.array_equal(m.numpy(), z["pad928_mask"])
    assert m[0, 0, 57].all() and not m[0, 0, :57].any()


def test_sine_encoding_channel_order_quirk():
    """positional_encoding.py:88-96 stacks on dim 4 of 5-D tensors: per axis the channels are
    [64 sines | 64 cosines], not interleaved."""
    mask = torch.zeros(1, 2, 3, 5, dtype=torch.bool)
    pe = po.sine_positional_encoding_3d(mask)
    assert pe.shape == (1, 2, 384, 3, 5)
    n_embed = torch.tensor([1.0, 2.0]) / (2 + 1e-6) * 2 * np.pi
    assert torch.allclose(pe[0, :, 0, 0, 0], torch.sin(n_embed), atol=1e-6)
    assert torch.allclose(pe[0, :, 64, 0, 0], torch.cos(n_embed), atol=1e-6)
