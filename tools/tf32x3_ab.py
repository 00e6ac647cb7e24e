"""A/B of the fp32-parity position MLP (192 -> 1024 -> 256 over the tokens of one 1600x640 sample): exact FFMA GEMMs
against 3xTF32 tcgen05 GEMMs -- time and error against float64 (see engine.TF32X3_FP32_PATH)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petr_b200 import ops                                # noqa: E402
from petr_b200.ops import ACT_RELU                       # noqa: E402

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
T = 24000
x = (torch.randn(T, 192, generator=g) * 4).to(dev)
w1, b1 = (torch.randn(1024, 192, generator=g) / 14).to(dev), torch.randn(1024, generator=g).to(dev)
w2, b2 = (torch.randn(256, 1024, generator=g) / 32).to(dev), torch.randn(256, generator=g).to(dev)
ref = (torch.relu(x.double() @ w1.double().T + b1.double()) @ w2.double().T + b2.double())
w1t, w2t = ops.x3_weight_tf32(w1), ops.x3_weight_tf32(w2)


def ffma():
    return ops.gemm_f32(ops.gemm_f32(x, w1, b1, act=ACT_RELU), w2, b2)


def tf32():
    h = ops.gemm_tf32x3(ops.split_rows_tf32(x), w1t, b1, act=ACT_RELU, split_out=True)
    return ops.gemm_tf32x3(h, w2t, b2)


for name, fn in (("FFMA", ffma), ("3xTF32 tcgen05", tf32)):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"{name:16s}: {e0.elapsed_time(e1) / 10:.3f} ms per sample, max err / scale = {err:.2e}")
