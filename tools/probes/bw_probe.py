"""HBM write / read / copy bandwidth with plain torch kernels (context for the store-bound GEMM epilogues)."""
import torch
dev = "cuda:0"
n = 1 << 30                                    # 1 Gi bf16 = 2 GB
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
ms = t(lambda: a.fill_(1.0)); print(f"fill  (write only) : {2 * n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: a.sum());      print(f"sum   (read only)  : {2 * n / ms / 1e6:8.1f} GB/s")
ms = t(lambda: b.copy_(a));   print(f"copy  (read+write) : {4 * n / ms / 1e6:8.1f} GB/s")
