"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, loaded through oracle/ref_loader.py) on the seeded cases of tests/cases.py.

Run in the build container only:   python -m oracle.make_golden
The fixtures are committed; the GPU box (no /root/reference) only reads them.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import petr_oracle as po          # noqa: E402
from oracle import ref_loader as rl           # noqa: E402
from tests import cases                       # noqa: E402


def build_reference(case):
    ns = rl.load_reference()
    cfg = cases.head_cfg(case)
    typ = cfg.pop("type")
    cfg["train_cfg"] = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        head = getattr(ns, typ)(**cfg)
    head.load_state_dict(cases.weights(case, head), strict=True)
    return head.eval()


def run_case(name):
    case = cases.CASES[name]
    wl, feats, metas = cases.inputs(case)
    head = build_reference(case)
    B, N, _, H, W = feats.shape
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = head([feats], metas)
        masks = po.key_padding_masks(metas, B, N, H, W)       # same ops as petr_head.py:383-394
        pos, cmask = head.position_embeding([feats], metas, masks)
    pos = pos.numpy()
    g = dict(all_cls_scores=out["all_cls_scores"].numpy(), all_bbox_preds=out["all_bbox_preds"].numpy(),
             coords_mask=cmask.numpy(), masks=masks.numpy(),
             pos_scale=np.float32(np.abs(pos).max()))
    if pos.size > 200000:
        g["pos_sample"] = pos.reshape(-1)[::cases.POS_SAMPLE_STRIDE].copy()
    else:
        g["coords_position_embeding"] = pos
    return g


def integer_fixtures():
    """Frustum grids / LID depth bins / pad masks produced by the reference's own code paths
    (position_embeding internals are re-evaluated with the reference's exact expressions via the
    loaded module: we call head.position_embeding with hooks on torch.meshgrid inputs)."""
    ns = rl.load_reference()
    out = {}
    shapes = {"800x320": (20, 50, 320, 800), "c5_1408x512": (16, 44, 512, 1408),
              "1600x640": (40, 100, 640, 1600), "pad928": (58, 100, 928, 1600)}
    for k, (H, W, ph, pw) in shapes.items():
        captured = {}
        orig = torch.meshgrid

        def spy(tensors, *a, **kw):
            captured["axes"] = [t.clone() for t in tensors]
            return orig(tensors, *a, **kw)

        head = rl.build_reference_head("projects/configs/petr/petr_vovnet_gridmask_p4_800x320.py")
        metas = [dict(pad_shape=[(ph, pw, 3)], img_shape=[(ph, pw, 3)],
                      lidar2img=[np.eye(4)])]
        feats = torch.zeros(1, 1, 256, H, W)
        torch.meshgrid = spy
        try:
            with torch.no_grad(), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                head.position_embeding([feats], metas, torch.zeros(1, 1, H, W, dtype=torch.bool))
        finally:
            torch.meshgrid = orig
        cw, ch, cd = captured["axes"]
        out[f"{k}_coords_w"] = cw.numpy()
        out[f"{k}_coords_h"] = ch.numpy()
        out[f"{k}_coords_d"] = cd.numpy()
    # padded-image key masks through the reference forward's own mask code (petr_head.py:383-394)
    import torch.nn.functional as F
    m = torch.ones(1, 1, 928, 1600)
    m[0, 0, :900, :1600] = 0
    out["pad928_mask"] = F.interpolate(m, size=(58, 100)).to(torch.bool).numpy()
    return out


def main():
    os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
    for name in cases.CASES:
        g = run_case(name)
        np.savez_compressed(cases.golden_path(name), **g)
        print(name, {k: getattr(v, "shape", None) for k, v in g.items()},
              os.path.getsize(cases.golden_path(name)) // 1024, "KiB")
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "integer_grids.npz"), **integer_fixtures())
    # state_dict key/shape listing of the reference heads (checkpoint contract)
    keys = {}
    for label, cfgp in [("PETRHead", "projects/configs/petr/petr_vovnet_gridmask_p4_800x320.py"),
                        ("PETRHead_c5", "projects/configs/petr/petr_r50dcn_gridmask_c5.py"),
                        ("PETRv2Head", "projects/configs/petrv2/petrv2_vovnet_gridmask_p4_800x320.py")]:
        h = rl.build_reference_head(cfgp)
        keys[label] = {k: list(v.shape) for k, v in h.state_dict().items()}
    with open(os.path.join(cases.GOLDEN_DIR, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("done")


if __name__ == "__main__":
    main()
