"""Is the K / V projection bound by its stores?  The same GEMM with K = 64 (one k-block: no MMA / operand time to speak of)
against K = 256, head-split and row-major stores, plus a plain device-to-device copy of the same number of bytes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petr_b200 import ops                                    # noqa: E402

dev = "cuda:0"
B, Lk = 16, 24000
M = B * Lk
g = torch.Generator().manual_seed(0)
kin = torch.randn(M, 256, generator=g).to(torch.bfloat16).to(dev)
wk = (torch.randn(1536, 256, generator=g) / 16).to(torch.bfloat16).to(dev)
bk = torch.randn(1536, generator=g).to(dev)


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out_hs = torch.empty((6, B, 8, Lk, 32), dtype=torch.bfloat16, device=dev)
out_rm = torch.empty((M, 1536), dtype=torch.bfloat16, device=dev)
src = torch.empty_like(out_rm)
gb = out_rm.numel() * 2 / 1e9
for K in (256, 64):
    a, w = kin[:, :K], wk[:, :K].contiguous()
    ms = t(lambda: ops.gemm_bf16_tc(a, w, bk, headsplit_Lk=Lk, out=out_hs))
    print(f"K = {K:3d} head-split: {ms:.4f} ms  ({gb / ms:.2f} TB/s of stores)")
    ms = t(lambda: ops.gemm_bf16_tc(a, w, bk, out=out_rm))
    print(f"K = {K:3d} row-major : {ms:.4f} ms  ({gb / ms:.2f} TB/s of stores)")
ms = t(lambda: out_rm.copy_(src))
print(f"torch copy of the same {gb:.2f} GB: {ms:.4f} ms ({gb / ms:.2f} TB/s written + as much read)")
ms = t(lambda: out_rm.zero_())
print(f"memset of {gb:.2f} GB: {ms:.4f} ms ({gb / ms:.2f} TB/s written)")
from petr_b200 import _lib                                   # noqa: E402
_lib.set_tuning("gemm_debug", 1)
for staged in (1, 0):
    _lib.set_tuning("gemm_staged", staged)
    ms = t(lambda: ops.gemm_bf16_tc(kin, wk, bk, headsplit_Lk=Lk, out=out_hs))
    print(f"K = 256 head-split, global stores SKIPPED (staged={staged}): {ms:.4f} ms")
_lib.set_tuning("gemm_debug", 2)
ms = t(lambda: ops.gemm_bf16_tc(kin, wk, bk, headsplit_Lk=Lk, out=out_hs))
print(f"K = 256 head-split, epilogue = TMEM loads only: {ms:.4f} ms")
_lib.set_tuning("gemm_debug", 0)
_lib.set_tuning("gemm_staged", 1)
