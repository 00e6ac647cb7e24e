#!/usr/bin/env python
"""Smallest end-to-end invocation of every kernel of libpetr_b200.so, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
Runs the padded / partly masked `tiny_padded` case (3 samples x 80 keys, 64 queries) through the fp32 and the bf16
(tcgen05) paths, the v2 head (`tiny_v2`), the key-split emulation and the NMS-free decode, and checks the results
against the committed golden outputs of the reference."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petr_b200 as pb  # noqa: E402
from tests import cases  # noqa: E402

DEV = "cuda:0"
for name in ("tiny_padded", "tiny_v2"):
    case = cases.CASES[name]
    wl, feats, metas = cases.inputs(case)
    head = pb.build_head(cases.head_cfg(case))
    head.load_state_dict(cases.weights(case, head), strict=True)
    head = head.eval().to(DEV)
    g = cases.load_golden(name)
    for prec, tol in (("fp32", 2e-3), ("bf16", 5e-2)):
        head.precision = prec
        out = head([feats.to(DEV)], metas)
        err = np.abs(out["all_bbox_preds"].cpu().numpy() - g["all_bbox_preds"]).max()
        print(f"{name} {prec}: max abs box err {err:.5f}")
        assert err <= tol
    if name == "tiny_padded":
        head.keysplit = dict(world=2, emulate=True)
        head([feats.to(DEV)], metas)
        head.keysplit = None
    res = head.get_bboxes(out, [dict() for _ in range(feats.shape[0])])
    print(f"{name}: decoded {[int(r[1].numel()) for r in res]} boxes per sample")
torch.cuda.synchronize()
print("sanitize_case ok")
