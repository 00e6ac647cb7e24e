"""One training step of the head at a named workload (for `ncu -k regex:attention_bwd ...` captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petr_b200 as pb                                   # noqa: E402
from petr_b200 import synth                              # noqa: E402
from tools.train_bench import gts                        # noqa: E402

wl = pb.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "vov-p4-1600x640"]
dev = torch.device("cuda", 0)
head = pb.build_head(wl.head_cfg())
head.load_state_dict(synth.head_state_dict(head, 0), strict=True)
head = head.to(dev).train()
metas = synth.make_img_metas(wl, 1, seed=0)
x = synth.make_features(wl, 1, seed=0).to(dev)
gt, lab = gts(1, 30, 1, dev)
out = head([x], metas)
sum(head.loss(gt, lab, out).values()).backward()
torch.cuda.synchronize()
print("step done")
