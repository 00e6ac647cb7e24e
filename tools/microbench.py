#!/usr/bin/env python
"""Kernel-level A/B timings on one B200 (CUDA events, inputs larger than L2 at the default sizes).

  python tools/microbench.py flash [--batch 16] [--keys 24000]
  python tools/microbench.py kproj [--batch 16]
  python tools/microbench.py k1    [--batch 16]

Prints one JSON line per variant.  Variants are selected through petr_set_tuning (explicit ABI)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petr_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def flash(args):
    B, Lq, Lk, H = args.batch, 900, args.keys, 8
    g = torch.Generator(device=DEV).manual_seed(0)
    q = torch.randn(B * Lq, 256, generator=g, device=DEV)
    Kh = torch.randn(B, H, Lk, 32, generator=g, device=DEV).to(torch.bfloat16)
    Vh = torch.randn(B, H, Lk, 32, generator=g, device=DEV).to(torch.bfloat16)
    qs = ops.pack_q_bf16(q, B, Lq, H)
    from petr_b200.engine_tc import choose_nsplit
    ns = args.nsplit or choose_nsplit(B, H, Lq, Lk)
    o_s = torch.empty((B * Lq, 512), dtype=torch.bfloat16, device=DEV)
    part = torch.empty((ns * ops.attention_partial_floats(B, Lq, H),), dtype=torch.float32, device=DEV) if ns > 1 else None
    ws = ops.cross_attention_workspace(B, Lq, Lk, H, ns, DEV)
    flops = 4.0 * Lq * Lk * 256 * B
    ref = None
    for mode in args.modes.split(","):
        if mode != "exact":
            _lib.set_tuning("flash_poly", int(mode[4:]))
        fn = lambda: ops.cross_attention_bf16(qs, Kh, Vh, None, B, Lq, Lk, H, ns, partial=part, out_split=o_s,
                                              want_f32=False, workspace=ws, exact_only=(mode == "exact"))
        med, best = timeit(fn)
        out = o_s[:, :256].float() + o_s[:, 256:].float()
        if ref is None:
            ref = out.clone()
        print(json.dumps(dict(kernel="flash", mode=mode, B=B, Lk=Lk, nsplit=ns, ms_median=med, ms_best=best,
                              tflops=flops / med / 1e9, flagged=int(ws.view(torch.int32).sum()) if mode != "exact" else None,
                              max_diff_vs_exact=float((out - ref).abs().max()))), flush=True)
    _lib.set_tuning("flash_poly", 2)


def kproj(args):
    B, Lk = args.batch, args.keys
    M, N, K = B * Lk, 1536, 256
    A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) / 16).to(torch.bfloat16)
    b = torch.randn(N, device=DEV)
    ref = None
    for staged, band in ((0, 0), (1, 0), (1, 1)):
        _lib.set_tuning("gemm_staged", staged)
        _lib.set_tuning("gemm_band", band)
        out = [None]

        def fn():
            out[0] = ops.gemm_bf16_tc(A, W, b, headsplit_Lk=Lk, name="gemm_kproj")
        med, best = timeit(fn)
        if ref is None:
            ref = out[0].clone()
        print(json.dumps(dict(kernel="kproj", staged=staged, band=band, M=M, ms_median=med, ms_best=best,
                              tflops=2.0 * M * N * K / med / 1e9, write_gbs=M * N * 2 / med / 1e6,
                              equal=bool(torch.equal(ref, out[0])))), flush=True)
    _lib.set_tuning("gemm_staged", 1)


def qside(args):
    """The query-side bf16x3 GEMM shapes of one decoder layer (M = B * 900 rows)."""
    M = args.batch * 900
    g = torch.Generator(device=DEV).manual_seed(1)
    X = torch.randn(M, 256, generator=g, device=DEV)
    Xs = ops.split_rows(X)
    H = torch.randn(M, 2048, generator=g, device=DEV)
    Hs = ops.split_rows(H)
    w256 = ops.x3_weight(torch.randn(256, 256, generator=g, device=DEV) / 16)
    w512 = ops.x3_weight(torch.randn(512, 256, generator=g, device=DEV) / 16)
    wf1 = ops.x3_weight(torch.randn(2048, 256, generator=g, device=DEV) / 16)
    wf2 = ops.x3_weight(torch.randn(256, 2048, generator=g, device=DEV) / 45)
    b256 = torch.randn(256, device=DEV); b512 = torch.randn(512, device=DEV); b2048 = torch.randn(2048, device=DEV)
    cases = {
        "proj256_res_f32": lambda: ops.gemm_bf16_tc(Xs, w256, b256, residual=X, x3=True),
        "qk512_headsplit_pair": lambda: ops.gemm_bf16_tc(Xs, w512, b512, x3=True, headsplit_Lk=900, headsplit_pair=True),
        "ffn1_relu_split": lambda: ops.gemm_bf16_tc(Xs, wf1, b2048, act=ops.ACT_RELU, x3=True, split_out=True),
        "ffn2_res_f32": lambda: ops.gemm_bf16_tc(Hs, wf2, b256, residual=X, x3=True),
    }
    for name, fn in cases.items():
        med, best = timeit(fn, iters=20)
        # the same launch 20 times inside one CUDA graph: per-kernel time without host launch overhead
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        gm, gb = timeit(g.replay, iters=5)
        print(json.dumps(dict(kernel="qside", case=name, M=M, us_eager_median=med * 1e3, us_in_graph=gm * 1e3 / 20)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["flash", "kproj", "qside", "all"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--keys", type=int, default=24000)
    ap.add_argument("--nsplit", type=int, default=0)
    ap.add_argument("--modes", default="exact,poly0,poly1,poly2,poly3,poly4")
    args = ap.parse_args()
    if args.what in ("flash", "all"):
        flash(args)
    if args.what in ("kproj", "all"):
        kproj(args)
    if args.what in ("qside", "all"):
        qside(args)


if __name__ == "__main__":
    main()
