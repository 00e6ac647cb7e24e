#!/usr/bin/env python
"""Timeline of one CTA of the flash cross-attention fast pass (clock64 stamps, petr_cross_attention_trace).
Prints, per key tile, when each role passed each pipeline event (clocks relative to the first stamp)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petr_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
B, Lq, Lk, H, ns = 16, 900, 24000, 8, 2
g = torch.Generator(device=DEV).manual_seed(0)
q = torch.randn(B * Lq, 256, generator=g, device=DEV)
Kh = torch.randn(B, H, Lk, 32, generator=g, device=DEV).to(torch.bfloat16)
Vh = torch.randn(B, H, Lk, 32, generator=g, device=DEV).to(torch.bfloat16)
qs = ops.pack_q_bf16(q, B, Lq, H)
part = torch.empty((ns * ops.attention_partial_floats(B, Lq, H),), dtype=torch.float32, device=DEV)
ws = ops.cross_attention_workspace(B, Lq, Lk, H, ns, DEV)
trace = torch.zeros((7, 64, 8), dtype=torch.int64, device=DEV)
for _ in range(3):
    _lib.call("petr_cross_attention_trace", _lib.ptr(qs), _lib.ptr(Kh), _lib.ptr(Vh), _lib.ptr(part), _lib.ptr(ws),
              _lib.ptr(trace), B, Lq, Lk, H, ns, _lib.stream_ptr(DEV))
torch.cuda.synchronize()
t = trace.cpu()
t0 = int(t[t > 0].min())
names = {0: "MMA-A", 1: "MMA-B", 2: "smA-h0", 3: "smA-h1", 4: "smB-h0", 5: "smB-h1", 6: "TMA"}
ev = {0: ["kv_full", "s_free->QK", "QK issued", "p_ready->PV", "PV issued"],
      2: ["s_full", "S loaded", "exps done", "pv_done ok", "P stored"], 6: ["kv_empty"]}
for j in list(range(0, 6)) + list(range(40, 46)):
    print(f"--- key tile {j}")
    for role in range(7):
        e = ev[0] if role < 2 else ev[6] if role == 6 else ev[2]
        vals = [int(t[role, j, k]) - t0 if t[role, j, k] > 0 else None for k in range(len(e))]
        print(f"  {names[role]:7s} " + "  ".join(f"{n}={v}" for n, v in zip(e, vals)))
# steady-state period per role
for role in (0, 1, 2, 4):
    k = 1 if role < 2 else 0
    d = [int(t[role, j + 1, k] - t[role, j, k]) for j in range(20, 60)]
    print(f"{names[role]}: mean period {sum(d) / len(d):.0f} clk (tiles 20-60), min {min(d)} max {max(d)}")
