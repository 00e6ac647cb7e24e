import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petr_b200 import ops
dev = "cuda:0"
B, Lk = 16, 24000
M = B * Lk
g = torch.Generator().manual_seed(0)
kin = torch.randn(M, 256, generator=g).to(torch.bfloat16).to(dev)
wk = (torch.randn(1536, 256, generator=g) / 16).to(torch.bfloat16).to(dev)
bk = torch.randn(1536, generator=g).to(dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out_hs = torch.empty((6, B, 8, Lk, 32), dtype=torch.bfloat16, device=dev)
out_rm = torch.empty((M, 1536), dtype=torch.bfloat16, device=dev)
print("head-split [layer][b][head][key][32]:", round(t(lambda: ops.gemm_bf16_tc(kin, wk, bk, headsplit_Lk=Lk, out=out_hs)), 4), "ms")
print("row-major  [token][layer*head*32]   :", round(t(lambda: ops.gemm_bf16_tc(kin, wk, bk, out=out_rm)), 4), "ms")
# per-layer token-major [layer][token][256]: six GEMMs N=256 each writing contiguous 64 KB tiles
outs = [torch.empty((M, 256), dtype=torch.bfloat16, device=dev) for _ in range(6)]
def six():
    for l in range(6):
        ops.gemm_bf16_tc(kin, wk[l*256:(l+1)*256], bk[l*256:(l+1)*256], out=outs[l])
print("6 x [token][256] (N=256 each)       :", round(t(six), 4), "ms")
