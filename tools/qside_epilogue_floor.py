"""Query-side GEMM shapes (B = 16 and 32: M = B * 900 rows) against their MMA-bound floor: the same launch with the
epilogue reduced to its TMEM loads (petr_set_tuning gemm_debug = 2).  Shows how much of each GEMM is epilogue."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petr_b200 import _lib, ops                              # noqa: E402
from petr_b200.ops import ACT_RELU                           # noqa: E402

dev = "cuda:0"


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for B in (16, 32):
    M = B * 900
    X = torch.randn(M, 256, device=dev)
    Xs = ops.split_rows(X)
    H = torch.randn(M, 2048, device=dev)
    Hs = ops.split_rows(H)
    w256 = ops.x3_weight(torch.randn(256, 256, device=dev) / 16)
    w512 = ops.x3_weight(torch.randn(512, 256, device=dev) / 16)
    w2048 = ops.x3_weight(torch.randn(2048, 256, device=dev) / 16)
    wf2 = ops.x3_weight(torch.randn(256, 2048, device=dev) / 45)
    b256, b512, b2048 = (torch.randn(n, device=dev) for n in (256, 512, 2048))
    cases = {
        "out-proj  N=256  K=3x256  fp32 + residual (mode 0)": lambda: ops.gemm_bf16_tc(Xs, w256, b256, residual=X, x3=True),
        "sa q,k    N=512  K=3x256  head-split pair (mode 4)": lambda: ops.gemm_bf16_tc(Xs, w512, b512, x3=True, headsplit_Lk=900, headsplit_pair=True),
        "FFN1      N=2048 K=3x256  relu, split out (mode 3)": lambda: ops.gemm_bf16_tc(Xs, w2048, b2048, act=ACT_RELU, x3=True, split_out=True),
        "FFN2      N=256  K=3x2048 fp32 + residual (mode 0)": lambda: ops.gemm_bf16_tc(Hs, wf2, b256, residual=X, x3=True),
    }
    for name, fn in cases.items():
        full = t(fn)
        _lib.set_tuning("gemm_debug", 2)
        floor = t(fn)
        _lib.set_tuning("gemm_debug", 0)
        print(f"B = {B:2d}  {name}: {full:7.1f} us, TMEM-loads-only epilogue {floor:7.1f} us")
