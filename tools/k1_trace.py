#!/usr/bin/env python
"""Timeline of CTA 0 of the fused position-embedding kernel (clock64 stamps via petr_debug_set_trace)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import petr_b200 as pb  # noqa: E402
from petr_b200 import _lib, ops, synth  # noqa: E402

DEV = "cuda:0"
wl = pb.WORKLOADS["vov-p4-1600x640"]
B = 16
metas = synth.make_img_metas(wl, B, 0)
from petr_b200 import engine  # noqa: E402
host = engine.host_prep(metas, wl.num_cams, wl.feat_h, wl.feat_w, False)
i2l = host["i2l"].to(DEV)
T = B * wl.num_keys
w1 = (torch.randn(1024, 192, device=DEV) * 0.05).to(torch.bfloat16)
w2 = (torch.randn(256, 1024, device=DEV) * 0.03).to(torch.bfloat16)
b1 = torch.randn(1024, device=DEV) * 0.05
b2 = torch.randn(256, device=DEV) * 0.05
mem = torch.randn(T, 256, device=DEV)
sin1 = torch.randn(wl.num_keys, 256, device=DEV)
out = torch.empty((T, 256), dtype=torch.bfloat16, device=DEV)
pr = (-61.2, -61.2, -10.0, 61.2, 61.2, 10.0)
trace = torch.zeros((4, 64, 8), dtype=torch.int64, device=DEV)
res = mem if "nores" not in sys.argv else None
for i in range(3):
    if i == 2:
        _lib.lib().petr_debug_set_trace(_lib.ptr(trace))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.position_mlp_fused(i2l, wl.feat_h, wl.feat_w, 64, 640, 1600, True, 1.0, (61.2 - 1.0) / (64 * 65), pr, w1, b1, w2, b2,
                           residual=res, residual2=sin1 if res is not None else None, want_f32=False, out_bf16=out)
    e1.record()
    torch.cuda.synchronize()
    print("ms", e0.elapsed_time(e1))
_lib.lib().petr_debug_set_trace(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
f = lambda v: (int(v) - t0) if v > 0 else None
print("MMA per chunk g: h_free ok | GEMM1 issued | p_ready(g-1) ok | GEMM2(g-1) issued ;  hidden-epi: h_full ok | done")
for g in list(range(0, 18)):
    print(f"g={g:2d}  MMA {f(t[0,g,0])} {f(t[0,g,1])} {f(t[0,g,2])} {f(t[0,g,3])}   HEPI {f(t[2,g,0])} {f(t[2,g,1])}")
print("tiles: producer a1_empty ok | done ;  output-epi: o_full ok | o_free | done")
for il in range(0, 5):
    print(f"tile {il}: PROD {f(t[1,il,0])} {f(t[1,il,1])}   OEPI {f(t[3,il,0])} {f(t[3,il,1])} {f(t[3,il,2])}")
