"""Times one training step of the head (forward in train mode + loss + backward) on the fp32 training path and prints
one JSON line with the per-kernel breakdown (CUDA events around every C-ABI call; the autograd thread's launches are
included because the events are recorded on the stream, not counted per thread).

    python tools/train_bench.py --workload vov-p4-800x320 --batch 2 --steps 5
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petr_b200 as pb                                   # noqa: E402
from petr_b200 import _lib, synth, train                 # noqa: E402


def gts(B, n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    out, lab = [], []
    for _ in range(B):
        ctr = torch.rand(n, 3, generator=g) * torch.tensor([80.0, 80.0, 6.0]) - torch.tensor([40.0, 40.0, 4.0])
        out.append(torch.cat([ctr, torch.rand(n, 3, generator=g) * 3 + 0.5, (torch.rand(n, 1, generator=g) - 0.5) * 6,
                              torch.randn(n, 2, generator=g)], 1).to(dev))
        lab.append(torch.randint(0, 10, (n,), generator=g).to(dev))
    return out, lab


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vov-p4-800x320")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--gts", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = pb.WORKLOADS[args.workload]
    head = pb.build_head(wl.head_cfg())
    head.load_state_dict(synth.head_state_dict(head, 0), strict=True)
    head = head.to(dev).train()
    B = args.batch
    metas = synth.make_img_metas(wl, B, seed=0)
    x = synth.make_features(wl, B, seed=0).to(dev)
    gt, lab = gts(B, args.gts, 1, dev)
    opt = torch.optim.AdamW([p for p in head.parameters() if p.requires_grad], lr=1e-5)

    def step():
        opt.zero_grad(set_to_none=True)
        out = head([x], metas)
        losses = head.loss(gt, lab, out)
        tot = sum(losses.values())
        tot.backward()
        opt.step()
        return tot

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tot = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    _lib.TIMERS = {}
    step()
    torch.cuda.synchronize()
    br = {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in _lib.TIMERS.items()}
    _lib.TIMERS = None
    br = dict(sorted(br.items(), key=lambda kv: -kv[1]))
    print(json.dumps(dict(what="training step (forward + loss + backward + AdamW), fp32 training path", workload=wl.name,
                          batch=B, num_keys=wl.num_keys, ms_per_step=round(ms, 2), samples_per_s=round(B * 1e3 / ms, 2),
                          loss=float(tot.detach()), kernel_ms_per_step=br,
                          peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))))


if __name__ == "__main__":
    main()
