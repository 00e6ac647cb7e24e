"""Diagnostics for the tcgen05 kernels (run on the GPU box): isolates GEMM layout, QK^T, PV and
softmax problems with structured inputs.  Prints a text report; not a test."""
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from petr_b200 import ops

DEV = "cuda:0"
torch.manual_seed(0)


def rep(name, got, ref):
    err = (got.float() - ref.float()).abs()
    print(f"{name}: max_err={err.max().item():.4e} mean_err={err.mean().item():.4e} "
          f"ref_max={ref.abs().max().item():.3f} nan={torch.isnan(got.float()).sum().item()}", flush=True)
    return err.max().item()


def gemm():
    for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 128, 256), (300, 200, 192)]:
        A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        W = torch.randn(N, K, device=DEV).to(torch.bfloat16) / math.sqrt(K)
        ref = A.float() @ W.float().T
        got = ops.gemm_bf16_tc(A, W)
        torch.cuda.synchronize()
        e = rep(f"gemm {M}x{N}x{K}", got, ref)
        if e > 1e-2:
            print(" got[0,:8]", got[0, :8].tolist()); print(" ref[0,:8]", ref[0, :8].tolist())
            print(" got[1,:8]", got[1, :8].tolist()); print(" ref[1,:8]", ref[1, :8].tolist())
            print(" got[64,:4]", got[64, :4].tolist(), "ref", ref[64, :4].tolist())


def attn():
    B, H = 1, 8
    for (Lq, Lk) in [(128, 128), (128, 256), (128, 1024), (300, 700)]:
        Kh = torch.randn(B, H, Lk, 32, device=DEV).to(torch.bfloat16)
        Vh = torch.randn(B, H, Lk, 32, device=DEV).to(torch.bfloat16)
        # 1. Q = 0 -> uniform attention: O = mean(V): isolates the P.V path (V descriptor, P layout)
        q0 = torch.zeros(B * Lq, 256, device=DEV)
        qs = ops.pack_q_bf16(q0, B, Lq, H)
        got = ops.cross_attention_bf16(qs, Kh, Vh, None, B, Lq, Lk, H, 1)
        torch.cuda.synchronize()
        ref = Vh.float().mean(2, keepdim=True).expand(B, H, Lq, 32).transpose(1, 2).reshape(B * Lq, 256)
        rep(f"attn uniform Lq={Lq} Lk={Lk}", got, ref)
        # 2. V = const per dim -> O = const: normalisation
        Vc = torch.arange(32, device=DEV).float().view(1, 1, 1, 32).expand(B, H, Lk, 32).contiguous().to(torch.bfloat16)
        q = torch.randn(B * Lq, 256, device=DEV)
        qs = ops.pack_q_bf16(q, B, Lq, H)
        got = ops.cross_attention_bf16(qs, Kh, Vc, None, B, Lq, Lk, H, 1)
        torch.cuda.synchronize()
        ref = torch.arange(32, device=DEV).float().repeat(H).view(1, 256).expand(B * Lq, 256)
        rep(f"attn constV  Lq={Lq} Lk={Lk}", got, ref)
        # 3. general
        s = qs.double() @ Kh.double().transpose(-1, -2) * math.log(2.0)
        ref = (torch.softmax(s, -1) @ Vh.double()).transpose(1, 2).reshape(B * Lq, 256).float()
        got = ops.cross_attention_bf16(qs, Kh, Vh, None, B, Lq, Lk, H, 1)
        torch.cuda.synchronize()
        e = rep(f"attn general Lq={Lq} Lk={Lk}", got, ref)
        if e > 2e-2:
            print(" got[0,:6]", got[0, :6].tolist()); print(" ref[0,:6]", ref[0, :6].tolist())
        # 4. sharp attention (one dominant key per query): S path + rescaling
        q = torch.randn(B * Lq, 256, device=DEV) * 10
        qs = ops.pack_q_bf16(q, B, Lq, H)
        s = qs.double() @ Kh.double().transpose(-1, -2) * math.log(2.0)
        ref = (torch.softmax(s, -1) @ Vh.double()).transpose(1, 2).reshape(B * Lq, 256).float()
        got = ops.cross_attention_bf16(qs, Kh, Vh, None, B, Lq, Lk, H, 1)
        torch.cuda.synchronize()
        rep(f"attn sharp   Lq={Lq} Lk={Lk}", got, ref)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("gemm", "all"):
        gemm()
    if which in ("attn", "all"):
        attn()
