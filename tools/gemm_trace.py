#!/usr/bin/env python
"""Timeline of CTA 0 of the tcgen05 GEMM on the K-projection shape (clock64 stamps via petr_debug_set_trace)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petr_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
B, Lk = 16, 24000
M, N, K = B * Lk, 1536, 256
A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
W = (torch.randn(N, K, device=DEV) / 16).to(torch.bfloat16)
b = torch.randn(N, device=DEV)
mode = sys.argv[1] if len(sys.argv) > 1 else "headsplit"
trace = torch.zeros((4, 64, 8), dtype=torch.int64, device=DEV)
if mode == "qside":
    M = 14400
    X = torch.randn(M, 256, device=DEV)
    A = ops.split_rows(X)
    W = ops.x3_weight(torch.randn(256, 256, device=DEV) / 16)
    b = torch.randn(256, device=DEV)
for i in range(3):
    if i == 2:
        _lib.lib().petr_debug_set_trace(_lib.ptr(trace))
    if mode == "qside":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_bf16_tc(A, W, b, residual=X, x3=True)
        e1.record()
        torch.cuda.synchronize()
        print("event time us", e0.elapsed_time(e1) * 1e3)
    elif mode == "headsplit":
        ops.gemm_bf16_tc(A, W, b, headsplit_Lk=Lk)
    elif mode == "bf16":
        ops.gemm_bf16_tc(A, W, b, out_dtype=torch.bfloat16)
    else:
        ops.gemm_bf16_tc(A, W, b)
torch.cuda.synchronize()
_lib.lib().petr_debug_set_trace(None)
t = trace.cpu()
t0 = int(t[t > 0].min())
names = ["TMA", "MMA", "EPI-lo", "EPI-hi"]
ev = [["kb0", "kb1", "kb2", "kb3"], ["acc_empty ok", "kb0 full", "kb1 full", "kb2 full", "kb3 full", "committed"],
      ["loop top", "acc_full ok", "chunk0", "chunk1", "chunk2", "chunk3", "arrived"]]
for j in ([0] if mode == "qside" else list(range(0, 3)) + list(range(30, 34))):
    print(f"--- tile {j}")
    for role in range(4):
        e = ev[min(role, 2)]
        print(f"  {names[role]:7s} " + "  ".join(f"{n}={int(t[role, j, k]) - t0 if t[role, j, k] > 0 else None}" for k, n in enumerate(e)))
if mode != "qside":
    for role, k in ((1, 5), (2, 6)):
        d = [int(t[role, j + 1, k] - t[role, j, k]) for j in range(10, 60)]
        print(f"{names[role]}: mean tile period {sum(d) / len(d):.0f} clk")
