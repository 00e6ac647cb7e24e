"""GPU parity tests of every C-ABI kernel against the CPU oracle / a plain fp32 evaluation.
Integer-class results (masks, frustum grid/depth/transform, coords_mask) must be bit-exact."""
import math

import numpy as np
import pytest
import torch

from oracle import petr_oracle as po
from petr_b200 import _lib, engine, ops, synth
from petr_b200.configs import WORKLOADS
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _metas_rand_shapes(wl, B, seed, pad):
    rng = np.random.RandomState(seed)
    metas = synth.make_img_metas(wl, B, seed, pad_shape=pad)
    for m in metas:
        m["img_shape"] = [(int(rng.randint(1, pad[0] + 1)), int(rng.randint(1, pad[1] + 1)), 3)
                          for _ in range(wl.num_cams)]
    return metas


@pytest.mark.parametrize("H,W,pad", [(4, 10, (64, 160)), (58, 100, (928, 1600)), (20, 50, (320, 800)),
                                     (16, 44, (512, 1408)), (7, 13, (100, 333)), (40, 100, (640, 1600))])
def test_key_padding_mask_bit_exact(H, W, pad):
    wl = WORKLOADS["tiny"]
    metas = _metas_rand_shapes(wl, 3, 5, pad)
    metas[0]["img_shape"] = [(pad[0], pad[1], 3)] * wl.num_cams
    metas[1]["img_shape"][0] = (0, 0, 3)
    ref = po.key_padding_masks(metas, 3, wl.num_cams, H, W)
    img_hw = torch.from_numpy(engine.img_hw_from_metas(metas, wl.num_cams)).to(DEV)
    got = ops.key_padding_mask(img_hw, pad[0], pad[1], H, W)
    assert torch.equal(got.cpu().bool(), ref)
    from petr_b200.engine_tc import any_key_masked
    assert any_key_masked(metas, wl.num_cams, H, W) == bool(ref.any())
    full = synth.make_img_metas(wl, 2, 0, img_shape=pad, pad_shape=pad)
    assert any_key_masked(full, wl.num_cams, H, W) is False


@pytest.mark.parametrize("wn,B,pr", [("tiny", 2, None), ("vov-p4-800x320", 1, None),
                                     ("tiny", 2, (-20.0, -20.0, -2.0, 20.0, 20.0, 2.0)),
                                     ("r50-c5-1408x512", 1, None)])
def test_frustum_bit_exact(wn, B, pr):
    wl = WORKLOADS[wn]
    pr = pr or (-61.2, -61.2, -10.0, 61.2, 61.2, 10.0)
    metas = synth.make_img_metas(wl, B, 7)
    N, H, W, D = wl.num_cams, wl.feat_h, wl.feat_w, 64
    masks = torch.zeros(B, N, H, W, dtype=torch.bool)
    masks[0, 0, 0, :3] = True
    feat_ref, cmask_ref, c3_ref = po.position_features(metas, B, N, H, W, masks, 64, 1.0, pr, True)
    i2l = torch.from_numpy(engine.img2lidar_from_metas(metas)).to(DEV)
    assert torch.equal(i2l.cpu(), po.img2lidar_matrices(metas))
    bin_size = (pr[3] - 1.0) / (64 * 65)
    feat, cmask, c3 = ops.frustum_features(i2l, masks.to(DEV).to(torch.uint8), H, W, D, wl.img_h, wl.img_w,
                                           True, 1.0, bin_size, pr, want_mask=True, want_coords=True)
    c3_ref = c3_ref.permute(0, 1, 3, 2, 4, 5).contiguous()            # (B,N,W,H,D,3) -> (B,N,H,W,D,3)
    assert torch.equal(c3.cpu(), c3_ref), "frustum transform / normalisation must be bit-exact"
    assert torch.equal(cmask.cpu().bool(), cmask_ref)
    fr = feat_ref.view(B * N, 3 * D, H * W).permute(0, 2, 1).reshape(-1, 3 * D)
    assert (feat.cpu() - fr).abs().max() <= 4e-6                      # logf vs torch.log: <= 2 ulp at 11.5
    featb, _, _ = ops.frustum_features(i2l, None, H, W, D, wl.img_h, wl.img_w, True, 1.0, bin_size, pr,
                                       dtype=torch.bfloat16, want_mask=False)
    assert (featb.float().cpu() - fr).abs().max() <= 0.05


def test_frustum_no_lid():
    wl = WORKLOADS["tiny"]
    metas = synth.make_img_metas(wl, 1, 3)
    pr = (-61.2, -61.2, -10.0, 61.2, 61.2, 10.0)
    _, cm_ref, c3_ref = po.position_features(metas, 1, 2, 4, 10, None, 16, 1.0, pr, False)
    i2l = torch.from_numpy(engine.img2lidar_from_metas(metas)).to(DEV)
    _, cm, c3 = ops.frustum_features(i2l, None, 4, 10, 16, 64, 160, False, 1.0, (pr[3] - 1.0) / 16, pr,
                                     want_mask=True, want_coords=True)
    assert torch.equal(c3.cpu(), c3_ref.permute(0, 1, 3, 2, 4, 5).contiguous())
    assert torch.equal(cm.cpu().bool(), cm_ref)


@pytest.mark.parametrize("shape", [(2, 2, 4, 10), (1, 6, 20, 50), (1, 12, 5, 7)])
@pytest.mark.parametrize("normalize", [True, False])
def test_sine3d_features(shape, normalize):
    g = torch.Generator().manual_seed(1)
    mask = torch.rand(shape, generator=g) < 0.2
    mask[0, 0] = False
    ref = po.sine_positional_encoding_3d(mask, 128, normalize=normalize)
    B, N, H, W = shape
    ref = ref.permute(0, 1, 3, 4, 2).reshape(-1, 384)
    dim_t = engine.sine_dim_t(128, 10000).to(DEV)
    got = ops.sine3d_features(mask.to(DEV).to(torch.uint8), dim_t, 128, normalize, 2 * math.pi, 1e-6, 0.0)
    assert (got.cpu() - ref).abs().max() <= (2e-6 if normalize else 2e-4)


def test_pos2posemb3d():
    ref_pts = torch.rand(900, 3, generator=torch.Generator().manual_seed(0))
    ref = po.pos2posemb3d(ref_pts)
    got = ops.pos2posemb3d(ref_pts.to(DEV), engine.sine_dim_t(128, 10000).to(DEV))
    assert (got.cpu() - ref).abs().max() <= 2e-6


@pytest.mark.parametrize("M,N,K", [(900, 256, 256), (128, 64, 16), (1000, 10, 256), (333, 1024, 192),
                                   (4224, 256, 2048), (77, 130, 100)])
def test_gemm_f32(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    A2 = torch.randn(7, K, generator=g).to(DEV)
    ref = (A.double() @ W.double().T + b.double()).float()
    assert (ops.gemm_f32(A, W, b) - ref).abs().max() <= 2e-5
    ref2 = torch.relu(((A + A2[torch.arange(M, device=DEV) % 7]).double() @ W.double().T + b.double())).float() + res
    out_b = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    got2 = ops.gemm_f32(A, W, b, A2=A2, residual=res, act=ops.ACT_RELU, out_bf16=out_b)
    assert (got2 - ref2).abs().max() <= 3e-5
    assert (out_b.float() - ref2).abs().max() <= 0.05
    gate = torch.randn(M, N, generator=g).to(DEV)
    ref3 = gate * torch.sigmoid(ref)
    assert (ops.gemm_f32(A, W, b, gate=gate, act=ops.ACT_SIGMOID_GATE) - ref3).abs().max() <= 2e-5
    # strided views (packed in-proj slices) and strided output columns
    big = torch.zeros(M, N + 6, device=DEV)
    ops.gemm_f32(A, W, b, out=big[:, 3:3 + N])
    assert (big[:, 3:3 + N] - ref).abs().max() <= 2e-5 and big[:, :3].abs().max() == 0


def test_gemm_f32_nchw():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 64, 4, 10, generator=g).to(DEV)
    w = torch.randn(256, 64, 1, 1, generator=g).to(DEV) * 0.1
    b = torch.randn(256, generator=g).to(DEV)
    ref = torch.nn.functional.conv2d(x.cpu(), w.cpu(), b.cpu()).permute(0, 2, 3, 1).reshape(-1, 256)
    got = ops.gemm_f32(x, w.reshape(256, 64).contiguous(), b, nchw_hw=40)
    assert (got.cpu() - ref).abs().max() <= 2e-5


def _attn_ref(Q, K, V, mask, B, Lq, Lk, heads):
    q = Q.double().view(B, Lq, heads, 32).transpose(1, 2)
    k = K.double().view(B, Lk, heads, 32).transpose(1, 2)
    v = V.double().view(B, Lk, heads, 32).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(32)
    if mask is not None:
        s = s.masked_fill(mask.bool()[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * Lq, heads * 32).float()


@pytest.mark.parametrize("B,Lq,Lk,scale", [(2, 64, 80, 1.0), (1, 900, 900, 1.0), (2, 100, 333, 6.0)])
def test_attention_f32(B, Lq, Lk, scale):
    g = torch.Generator().manual_seed(Lk)
    Q = (torch.randn(B * Lq, 256, generator=g) * scale).to(DEV)
    K = (torch.randn(B * Lk, 256, generator=g) * scale).to(DEV)
    V = torch.randn(B * Lk, 256, generator=g).to(DEV)
    mask = (torch.rand(B, Lk, generator=g) < 0.3).to(DEV).to(torch.uint8)
    mask[0, :70] = 1
    for m in (None, mask):
        got = ops.attention_f32(Q, K, V, m, B, Lq, Lk, 8)
        assert (got - _attn_ref(Q, K, V, m, B, Lq, Lk, 8)).abs().max() <= 1e-4
    mask[B - 1] = 1                                               # fully masked sample -> NaN like torch
    got = ops.attention_f32(Q, K, V, mask, B, Lq, Lk, 8)
    assert torch.isnan(got[(B - 1) * Lq:]).all()


def test_layernorm_and_decode():
    g = torch.Generator().manual_seed(0)
    X = torch.randn(1000, 256, generator=g).to(DEV) * 3 + 1
    gm, bt = torch.randn(256, generator=g).to(DEV), torch.randn(256, generator=g).to(DEV)
    ref = torch.nn.functional.layer_norm(X, (256,), gm, bt, 1e-5)
    assert (ops.layernorm(X, gm, bt) - ref).abs().max() <= 1e-5
    assert (ops.layernorm(X, gm, bt, relu=True) - torch.relu(ref)).abs().max() <= 1e-5
    Xn = X.clone()
    Xn[5] = float("nan")
    out = ops.layernorm(Xn, gm, bt, nan_to_num=True)
    assert out[5].abs().max() == 0 and (out[6] - ref[6]).abs().max() <= 1e-5
    L, B, Q = 2, 3, 50
    reg = torch.randn(L * B * Q, 10, generator=g).to(DEV)
    refp = torch.rand(Q, 3, generator=g).to(DEV)
    td = torch.tensor([0.5, 0.25, 1.0], device=DEV)
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    exp = reg.clone().view(L, B, Q, 10)
    r = po.inverse_sigmoid(refp.cpu()).to(DEV)
    exp[..., 0:2] = (exp[..., 0:2] + r[:, 0:2]).sigmoid()
    exp[..., 4:5] = (exp[..., 4:5] + r[:, 2:3]).sigmoid()
    exp[..., 8:] = exp[..., 8:] / td[None, :, None, None]
    exp[..., 0] = exp[..., 0] * 102.4 - 51.2
    exp[..., 1] = exp[..., 1] * 102.4 - 51.2
    exp[..., 4] = exp[..., 4] * 8.0 - 5.0
    got = ops.decode_boxes(reg.clone(), refp, pc, td, L, B, Q).view(L, B, Q, 10)
    assert (got - exp).abs().max() <= 2e-5


# ---------------------------------------------------------------------------------------------
# tcgen05 tier
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (1000, 1024, 192), (4224, 256, 1024),
                                   (333, 1536, 256), (130, 100, 384)])
def test_gemm_bf16_tc(M, N, K):
    g = torch.Generator().manual_seed(M * 3 + N + K)
    A = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).to(torch.bfloat16)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = (A.double() @ W.double().T + b.double()).float()
    got = ops.gemm_bf16_tc(A, W, b)
    assert (got - ref).abs().max() <= 2e-4, "tcgen05 GEMM (fp32 out)"
    got = ops.gemm_bf16_tc(A, W, b, residual=res, act=ops.ACT_RELU)
    assert (got - (torch.relu(ref) + res)).abs().max() <= 2e-4
    gb = ops.gemm_bf16_tc(A, W, b, act=ops.ACT_RELU, out_dtype=torch.bfloat16)
    assert (gb.float() - torch.relu(ref)).abs().max() <= 0.05


@pytest.mark.parametrize("M,N,K0", [(900, 256, 256), (3600, 2048, 256), (1000, 256, 2048), (130, 512, 192)])
def test_gemm_bf16x3_fp32_class(M, N, K0):
    """A = [hi | lo], W = [hi | hi | lo]: the bf16x3 product must be fp32-class (~2^-16)."""
    g = torch.Generator().manual_seed(M + N + K0)
    A = torch.randn(M, K0, generator=g).to(DEV)
    W = (torch.randn(N, K0, generator=g) / math.sqrt(K0)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = (A.double() @ W.double().T + b.double()).float()
    As = ops.split_rows(A)
    assert (As[:, :K0].float() + As[:, K0:].float() - A).abs().max() <= 2e-5 * A.abs().max()
    W3 = ops.x3_weight(W)
    got = ops.gemm_bf16_tc(As, W3, b, x3=True)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max() <= 4e-5 * scale, (got - ref).abs().max()
    # relu + residual + split output with fp32 copy
    f32 = torch.empty(M, N, device=DEV)
    sp = ops.gemm_bf16_tc(As, W3, b, x3=True, act=ops.ACT_RELU, residual=res, split_out=True, out_f32=f32)
    ref2 = torch.relu(ref) + res
    assert (f32 - ref2).abs().max() <= 4e-5 * scale
    assert (sp[:, :N].float() + sp[:, N:].float() - ref2).abs().max() <= 6e-5 * scale
    # alpha + head-split (the Q projection form)
    if N == 256 and M % 100 == 0:
        hs = ops.gemm_bf16_tc(As, W3, b, x3=True, alpha=0.25, headsplit_Lk=100)
        refh = (0.25 * ref).view(M // 100, 100, 8, 32).permute(0, 2, 1, 3)
        assert (hs[0].float() - refh).abs().max() <= 0.02 * max(1.0, 0.25 * scale)


def test_layernorm_multi_split_rows_nchw():
    g = torch.Generator().manual_seed(2)
    R, E, Lq = 700, 256, 100
    X = (torch.randn(R, E, generator=g) * 2 + 0.5).to(DEV)
    qpos = torch.randn(Lq, E, generator=g).to(DEV)
    g1, b1, g2, b2 = (torch.randn(E, generator=g).to(DEV) for _ in range(4))
    Y = torch.empty(R, E, device=DEV); Z = torch.empty(R, E, device=DEV)
    Ys, Yps, Zs = (torch.empty(R, 2 * E, dtype=torch.bfloat16, device=DEV) for _ in range(3))
    Xn = X.clone(); Xn[3] = float("nan")
    ops.layernorm_multi(Xn, g1, b1, qpos=qpos, Y=Y, Y_split=Ys, Ypos_split=Yps, gamma2=g2, beta2=b2, Z=Z, Z_split=Zs)
    yref = torch.nn.functional.layer_norm(X, (E,), g1, b1, 1e-5)
    zref = torch.nn.functional.layer_norm(yref, (E,), g2, b2, 1e-5)
    ok = torch.ones(R, dtype=torch.bool, device=DEV); ok[3] = False
    assert (Y[ok] - yref[ok]).abs().max() <= 2e-5 and (Z[ok] - zref[ok]).abs().max() <= 5e-5
    assert torch.isnan(Y[3]).all() and Z[3].abs().max() == 0 and Zs[3].float().abs().max() == 0
    yp = yref + qpos[torch.arange(R, device=DEV) % Lq]
    for sp, r in ((Ys, yref), (Yps, yp), (Zs, zref)):
        assert (sp[ok][:, :E].float() + sp[ok][:, E:].float() - r[ok]).abs().max() <= 1e-4
    ops.layernorm_multi(X, g1, b1, Y=Y, relu=True)
    assert (Y - torch.relu(yref)).abs().max() <= 2e-5
    sp = ops.split_rows(X, add=qpos, relu=True)
    r = torch.relu(X + qpos[torch.arange(R, device=DEV) % Lq])
    assert (sp[:, :E].float() + sp[:, E:].float() - r).abs().max() <= 1e-4
    x = torch.randn(5, 70, 6, 11, generator=g).to(DEV)
    t = ops.nchw_to_tokens_bf16(x)
    assert torch.equal(t, x.permute(0, 2, 3, 1).reshape(-1, 70).to(torch.bfloat16))


def test_attention_split_outputs():
    B, Lq, Lk = 2, 100, 260
    g = torch.Generator().manual_seed(4)
    Q = torch.randn(B * Lq, 256, generator=g).to(DEV)
    K = torch.randn(B * Lk, 256, generator=g).to(DEV)
    V = torch.randn(B * Lk, 256, generator=g).to(DEV)
    ref = _attn_ref(Q, K, V, None, B, Lq, Lk, 8)
    sp = torch.empty(B * Lq, 512, dtype=torch.bfloat16, device=DEV)
    ops.attention_f32(Q, K, V, None, B, Lq, Lk, 8, out_split=sp, want_f32=False)
    assert (sp[:, :256].float() + sp[:, 256:].float() - ref).abs().max() <= 1e-4
    qs, Kh, Vh, s = _flash_inputs(B, Lq, Lk, 2.0, 3)
    ref = (torch.softmax(s, -1) @ Vh.double()).transpose(1, 2).reshape(B * Lq, 256).float()
    for ns in (1, 2):
        f32 = torch.empty(B * Lq, 256, device=DEV)
        ops.cross_attention_bf16(qs, Kh, Vh, None, B, Lq, Lk, 8, ns, out=f32, out_split=sp)
        assert (f32 - ref).abs().max() <= 1e-2
        assert (sp[:, :256].float() + sp[:, 256:].float() - f32).abs().max() <= 1e-4


@pytest.mark.parametrize("B,Lq,Lk,scale,nsplit", [(2, 900, 900, 1.0, 1), (1, 128, 128, 1.0, 1), (3, 100, 333, 6.0, 1),
                                                  (1, 300, 1500, 3.0, 3)])
def test_attention_x3_tc_fp32_class(B, Lq, Lk, scale, nsplit):
    """hi/lo split attention on tcgen05 must match an fp64 evaluation to fp32-class accuracy, incl.
    through the head-split pair epilogue of the bf16x3 GEMM that feeds it."""
    g = torch.Generator().manual_seed(Lk + 7)
    E = 256
    X = (torch.randn(B * Lq, E, generator=g) * scale).to(DEV)
    Y = (torch.randn(B * Lk, E, generator=g) * scale).to(DEV)
    Wqk = (torch.randn(2 * E, E, generator=g) / 16).to(DEV)
    Wv = (torch.randn(E, E, generator=g) / 16).to(DEV)
    bqk = torch.randn(2 * E, generator=g).to(DEV)
    if Lq == Lk:
        Y = X
    q = (X.double() @ Wqk[:E].double().T + bqk[:E].double()).float()
    k = (Y.double() @ Wqk[E:].double().T + bqk[E:].double()).float()
    v = (Y.double() @ Wv.double().T).float()
    mask = (torch.rand(B, Lk, generator=g) < 0.2).to(DEV).to(torch.uint8) if Lk == 333 else None
    ref = _attn_ref(q, k, v, mask, B, Lq, Lk, 8)
    qp = ops.gemm_bf16_tc(ops.split_rows(X), ops.x3_weight(Wqk[:E].contiguous()), bqk[:E].contiguous(), x3=True,
                          headsplit_Lk=Lq, headsplit_pair=True)
    kp = ops.gemm_bf16_tc(ops.split_rows(Y), ops.x3_weight(Wqk[E:].contiguous()), bqk[E:].contiguous(), x3=True,
                          headsplit_Lk=Lk, headsplit_pair=True)
    vp = ops.gemm_bf16_tc(ops.split_rows(Y), ops.x3_weight(Wv), None, x3=True, headsplit_Lk=Lk, headsplit_pair=True)
    qh = (qp[0, 0].float() + qp[1, 0].float()).permute(0, 2, 1, 3).reshape(B * Lq, E)
    assert (qh - q).abs().max() <= 1e-4 * max(1.0, q.abs().max().item())
    sp = torch.empty(B * Lq, 2 * E, dtype=torch.bfloat16, device=DEV)
    for exact_only in (False, True):                 # max-free fast pass (+ exact pass on flagged CTAs) and exact alone
        got = ops.attention_x3_tc(qp[:, 0], kp[:, 0], vp[:, 0], mask, B, Lq, Lk, 8, nsplit, out_split=sp,
                                  exact_only=exact_only)
        err = (got - ref).abs().max().item()
        assert err <= 3e-4 * max(1.0, ref.abs().max().item()), (exact_only, err)
        assert (sp[:, :E].float() + sp[:, E:].float() - got).abs().max() <= 2e-5 * max(1.0, got.abs().max().item())


def test_gemm_bf16_tc_headsplit():
    B, Lk, K = 2, 200, 256
    g = torch.Generator().manual_seed(9)
    A = torch.randn(B * Lk, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(512, K, generator=g) / 16).to(DEV).to(torch.bfloat16)
    b = torch.randn(512, generator=g).to(DEV)
    ref = (A.double() @ W.double().T + b.double()).float()                # (B*Lk, 512)
    ref = ref.view(B, Lk, 2, 8, 32).permute(2, 0, 3, 1, 4)                # (L,B,heads,Lk,32)
    got = ops.gemm_bf16_tc(A, W, b, headsplit_Lk=Lk)
    assert got.shape == (2, B, 8, Lk, 32)
    assert (got.float() - ref).abs().max() <= 0.05


@pytest.mark.parametrize("M,N", [(148 * 128, 256), (19000, 256), (40 * 128 + 5, 1536), (2 * 24000, 1536)])
def test_gemm_bf16_tc_large_m(M, N):
    """K = 256 with more output tiles than SMs (the key-side GEMMs over B * Lk tokens): several
    tiles per persistent CTA, bias tiles prefetched one tile ahead, all output forms, ragged last
    m-tile, n-tile changing from one tile of a CTA to the next."""
    K = 256
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / 16).to(DEV).to(torch.bfloat16)
    b = torch.randn(N, generator=g).to(DEV)
    ref = A.float() @ W.float().T + b                                      # fp32 on bf16-exact inputs
    got = ops.gemm_bf16_tc(A, W, b)
    assert (got - ref).abs().max() <= 3e-4
    f32 = torch.empty(M, N, device=DEV)
    sp = ops.gemm_bf16_tc(A, W, b, split_out=True, out_f32=f32)
    assert (f32 - ref).abs().max() <= 3e-4
    assert (sp[:, :N].float() + sp[:, N:].float() - ref).abs().max() <= 6e-5 * ref.abs().max()
    if M % 24000 == 0:
        Lk = 24000
        hs = ops.gemm_bf16_tc(A, W, b, headsplit_Lk=Lk)
        want = ref.view(M // Lk, Lk, N // 256, 8, 32).permute(2, 0, 3, 1, 4)
        assert hs.shape == want.shape
        assert (hs.float() - want).abs().max() <= 0.05


def _flash_inputs(B, Lq, Lk, qscale, seed):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B * Lq, 256, generator=g) * qscale).to(DEV)
    Kh = torch.randn(B, 8, Lk, 32, generator=g).to(DEV).to(torch.bfloat16)
    Vh = torch.randn(B, 8, Lk, 32, generator=g).to(DEV).to(torch.bfloat16)
    qs = ops.pack_q_bf16(q, B, Lq, 8)
    # fp64 evaluation on the bf16-rounded operands (scale already folded into qs, exp2 domain)
    s = qs.double() @ Kh.double().transpose(-1, -2) * math.log(2.0)
    return qs, Kh, Vh, s


@pytest.mark.parametrize("mode", ["exact", "poly0", "poly1", "poly2", "poly3", "poly4"])
@pytest.mark.parametrize("B,Lq,Lk,qscale,nsplit", [
    (1, 128, 128, 1.0, 1), (1, 128, 256, 1.0, 1), (2, 900, 1000, 1.0, 1), (1, 900, 6000, 1.0, 1),
    (1, 900, 6000, 1.0, 3), (2, 64, 333, 8.0, 2), (1, 900, 4224, 12.0, 1), (1, 200, 1500, 1.0, 5),
    (1, 260, 700, 40.0, 2), (1, 33, 130, 1.0, 1)])
def test_cross_attention_bf16(B, Lq, Lk, qscale, nsplit, mode):
    """Both passes of the flash kernel: the max-free fast pass (with 0..4 of every 4 exponential pairs on the FMA
    pipe) + exact pass on the CTAs it flags, and the exact online softmax alone.  qscale 40 pushes
    row sums far outside the fast pass's window, so every CTA must fall back there."""
    qs, Kh, Vh, s = _flash_inputs(B, Lq, Lk, qscale, Lk + nsplit)
    ref = (torch.softmax(s, -1) @ Vh.double()).transpose(1, 2).reshape(B * Lq, 256).float()
    ws = ops.cross_attention_workspace(B, Lq, Lk, 8, nsplit, DEV)
    ws.fill_(0xFF)
    default_poly = _lib.get_tuning("flash_poly")
    _lib.set_tuning("flash_poly", int(mode[4:]) if mode.startswith("poly") else default_poly)
    try:
        got = ops.cross_attention_bf16(qs, Kh, Vh, None, B, Lq, Lk, 8, nsplit, workspace=ws,
                                       exact_only=(mode == "exact"))
    finally:
        _lib.set_tuning("flash_poly", default_poly)
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"flash attention error {err}"
    assert err <= 1.5e-2, f"flash attention error {err}"
    if mode != "exact":
        flags = ws.view(torch.int32).cpu()
        assert ((flags == 0) | (flags == 1)).all()
        if qscale <= 1.0:
            assert int(flags.sum()) == 0, "fast pass must hold for O(1) logits"
        if qscale >= 40.0:
            assert int(flags.sum()) == flags.numel(), "extreme logits must fall back to the exact pass"


def test_cross_attention_bf16_masks_and_merge():
    B, Lq, Lk = 3, 100, 700
    qs, Kh, Vh, s = _flash_inputs(B, Lq, Lk, 4.0, 5)
    g = torch.Generator().manual_seed(1)
    mask = (torch.rand(B, Lk, generator=g) < 0.4).to(DEV)
    mask[0, 128:384] = True                      # two fully masked key tiles
    mask[0, :128] = True                         # ... including the first one
    mask[2] = True                               # fully masked sample
    sm = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sm, -1) @ Vh.double()).transpose(1, 2).reshape(B * Lq, 256).float()
    m8 = mask.to(torch.uint8)
    for ns in (1, 3):
        got = ops.cross_attention_bf16(qs, Kh, Vh, m8, B, Lq, Lk, 8, ns)
        assert torch.isnan(got[2 * Lq:]).all()
        assert (got[:2 * Lq] - ref[:2 * Lq]).abs().max() <= 1e-2
    # key-split across "ranks": two halves computed separately, partials concatenated, merged
    half = 384
    parts = []
    for lo, hi in ((0, half), (half, Lk)):
        parts.append(ops.cross_attention_bf16(qs, Kh[:, :, lo:hi].contiguous(), Vh[:, :, lo:hi].contiguous(),
                                              m8[:, lo:hi].contiguous(), B, Lq, hi - lo, 8, 1,
                                              write_partial=True))
    merged = ops.attention_merge(torch.cat(parts), 2, B, Lq, 8)
    assert (merged[:2 * Lq] - ref[:2 * Lq]).abs().max() <= 1e-2
    assert torch.isnan(merged[2 * Lq:]).all()


@pytest.mark.parametrize("wn,B", [("tiny", 2), ("small", 3), ("vov-p4-800x320", 1)])
def test_position_mlp_fused(wn, B):
    """K1: the fused frustum -> inverse_sigmoid -> tcgen05 MLP kernel against (a) the unfused bf16
    path (same bf16 operands; only summation order / fast-math coordinate arithmetic differ) and
    (b) the fp32 oracle to bf16-class accuracy."""
    wl = WORKLOADS[wn]
    g = torch.Generator().manual_seed(17)
    metas = synth.make_img_metas(wl, B, 5)
    N, H, W = wl.num_cams, wl.feat_h, wl.feat_w
    pr = (-61.2, -61.2, -10.0, 61.2, 61.2, 10.0)
    bin_size = (pr[3] - 1.0) / (64 * 65)
    w1 = (torch.randn(1024, 192, generator=g) * 0.05).to(DEV)
    b1 = (torch.randn(1024, generator=g) * 0.05).to(DEV)
    w2 = (torch.randn(256, 1024, generator=g) * 0.03).to(DEV)
    b2 = (torch.randn(256, generator=g) * 0.05).to(DEV)
    T = B * N * H * W
    res = torch.randn(T, 256, generator=g).to(DEV)
    i2l = torch.from_numpy(engine.img2lidar_from_metas(metas)).to(DEV)
    w1b, w2b = w1.to(torch.bfloat16).contiguous(), w2.to(torch.bfloat16).contiguous()
    got = ops.position_mlp_fused(i2l, H, W, 64, wl.img_h, wl.img_w, True, 1.0, bin_size, pr, w1b, b1, w2b, b2,
                                 residual=res)
    feat, _, _ = ops.frustum_features(i2l, None, H, W, 64, wl.img_h, wl.img_w, True, 1.0, bin_size, pr,
                                      dtype=torch.bfloat16, want_mask=False)
    h1 = ops.gemm_bf16_tc(feat, w1b, b1, act=ops.ACT_RELU, out_dtype=torch.bfloat16)
    unf = ops.gemm_bf16_tc(h1, w2b, b2, residual=res)
    scale = (unf - res).abs().max().item()
    assert (got - unf).abs().max() <= 1e-2 * scale, ((got - unf).abs().max().item(), scale)
    feat32, _, _ = po.position_features(metas, B, N, H, W, None, 64, 1.0, pr, True)
    f = feat32.view(B * N, 192, H * W).permute(0, 2, 1).reshape(-1, 192).to(DEV)
    ref = torch.relu(f @ w1.T + b1) @ w2.T + b2 + res
    assert (got - ref).abs().max() <= 2e-2 * scale
    got0 = ops.position_mlp_fused(i2l, H, W, 64, wl.img_h, wl.img_w, True, 1.0, bin_size, pr, w1b, b1, w2b, b2)
    assert (got0 + res - got).abs().max() <= 1e-5 * max(1.0, scale)


def test_position_mlp_fused_broadcast_residual_and_add_rows():
    """K1 with the batch-invariant second residual + bf16 store (the direct K-projection input), and
    the add_rows_bf16 helper of the unfused / fpe variants."""
    wl = WORKLOADS["small"]
    B, N, H, W = 3, wl.num_cams, wl.feat_h, wl.feat_w
    g = torch.Generator().manual_seed(23)
    metas = synth.make_img_metas(wl, B, 9)
    pr = (-61.2, -61.2, -10.0, 61.2, 61.2, 10.0)
    bin_size = (pr[3] - 1.0) / (64 * 65)
    w1 = (torch.randn(1024, 192, generator=g) * 0.05).to(DEV).to(torch.bfloat16)
    b1 = (torch.randn(1024, generator=g) * 0.05).to(DEV)
    w2 = (torch.randn(256, 1024, generator=g) * 0.03).to(DEV).to(torch.bfloat16)
    b2 = (torch.randn(256, generator=g) * 0.05).to(DEV)
    Lk = N * H * W
    res = torch.randn(B * Lk, 256, generator=g).to(DEV)
    res2 = torch.randn(Lk, 256, generator=g).to(DEV)
    i2l = torch.from_numpy(engine.img2lidar_from_metas(metas)).to(DEV)
    args = (i2l, H, W, 64, wl.img_h, wl.img_w, True, 1.0, bin_size, pr, w1, b1, w2, b2)
    base = ops.position_mlp_fused(*args, residual=res)
    f32 = torch.empty_like(base)
    b16 = torch.empty(B * Lk, 256, dtype=torch.bfloat16, device=DEV)
    ops.position_mlp_fused(*args, residual=res, residual2=res2, out=f32, out_bf16=b16)
    exp = base + res2.repeat(B, 1)
    assert (f32 - exp).abs().max() <= 1e-5 * max(1.0, exp.abs().max().item())
    assert torch.equal(b16, f32.to(torch.bfloat16))
    assert torch.equal(ops.add_rows_bf16(base, res2), exp.to(torch.bfloat16))


@pytest.mark.parametrize("B,Q,C,code,max_num,thr", [(3, 900, 10, 10, 300, None), (2, 900, 10, 10, 300, 0.2),
                                                   (1, 64, 10, 10, 100, None), (2, 100, 10, 8, 300, 0.05),
                                                   (1, 30, 3, 10, 100, None)])
def test_nms_free_decode_kernel(B, Q, C, code, max_num, thr):
    """NMSFreeCoder.decode_single + get_bboxes' bottom-centre shift as ONE kernel launch per batch
    (nms_free_coder.py:48-119, petr_head.py:731-751) against the CPU oracle: labels and query indices bit-exact,
    scores / boxes to fp32 rounding of exp / atan2; also through the module API and inside a CUDA graph."""
    import petr_b200 as pb
    g = torch.Generator().manual_seed(B * 1000 + Q)
    cls = torch.randn(B, Q, C, generator=g) - 2.0
    box = torch.randn(B, Q, code, generator=g)
    box[..., 0:2] *= 40.0
    rng = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    kmax = min(max_num, Q * C)
    for bottom in (False, True):
        ref = po.nms_free_decode(cls, box, kmax, C, rng, thr, bottom_center=bottom)
        boxes, scores, labels, index, mask, count = ops.nms_free_decode(cls.to(DEV), box.to(DEV), kmax, rng, thr, bottom)
        assert boxes.shape == (B, kmax, 9 if code > 8 else 7)
        for b in range(B):
            m = mask[b].cpu()
            assert int(count[b]) == int(m.sum()) == ref[b]["labels"].numel()
            assert torch.equal(labels[b].cpu()[m], ref[b]["labels"]), "labels must be bit-exact"
            assert torch.equal(index[b].cpu()[m], ref[b]["index"]), "query indices must be bit-exact"
            assert (scores[b].cpu()[m] - ref[b]["scores"]).abs().max() <= 2e-7
            assert (scores[b, :-1] >= scores[b, 1:]).all()
            d = (boxes[b].cpu()[m] - ref[b]["bboxes"]).abs()
            assert (d <= 1e-5 * ref[b]["bboxes"].abs().clamp(min=1.0)).all()
    if code == 10 and C == 10:
        coder = pb.NMSFreeCoder(pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], post_center_range=rng,
                                max_num=kmax, score_threshold=thr, num_classes=C)
        outs = dict(all_cls_scores=cls[None].to(DEV), all_bbox_preds=box[None].to(DEV))
        dec = coder.decode(outs)
        ref = po.nms_free_decode(cls, box, kmax, C, rng, thr)
        for d, r in zip(dec, ref):
            assert torch.equal(d["labels"].cpu(), r["labels"]) and d["bboxes"].shape == r["bboxes"].shape
        # graph capture: features -> boxes without leaving the stream
        xs, bs = cls.to(DEV), box.to(DEV)
        coder.decode_padded(xs, bs)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = coder.decode_padded(xs, bs, bottom_center=True)
        gr.replay()
        torch.cuda.synchronize()
        eager = coder.decode_padded(xs, bs, bottom_center=True)
        for a, b_ in zip(out, eager):
            assert torch.equal(a, b_)


@pytest.mark.parametrize("M,K,N", [(333, 192, 1024), (4096 + 77, 1024, 256), (128, 384, 100), (1000, 256, 260)])
def test_gemm_tf32x3_fp32_class(M, K, N):
    """3xTF32 tcgen05 GEMM (fp32 operands split hi/lo, kind::tf32): fp32-class accuracy against float64 (measured
    2e-6 of the output scale: the operands carry 22 mantissa bits, the remainder is the tensor core's fp32 accumulation),
    bias / ReLU / residual epilogues, and the [hi | lo] split output feeding the next GEMM (the fp32-parity position MLP)."""
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g) * 3
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    R = torch.randn(M, N, generator=g)
    ref = A.double() @ W.double().T + b.double()
    As = ops.split_rows_tf32(A.to(DEV))
    assert torch.equal(As[:, :K] + As[:, K:], A.to(DEV))                       # hi + lo == value exactly
    W3 = ops.x3_weight_tf32(W.to(DEV))
    out = ops.gemm_tf32x3(As, W3, b.to(DEV))
    scale = ref.abs().max().item()
    err = (out.cpu().double() - ref).abs().max().item()
    print(f"3xTF32 {M}x{K}x{N}: err / scale = {err / scale:.2e}")
    tol = 2e-6 + 1e-8 * K          # the tensor core's truncating fp32 accumulator: error grows with the K steps
    assert err <= tol * scale, (err, scale)
    out = ops.gemm_tf32x3(As, W3, b.to(DEV), residual=R.to(DEV), act=ops.ACT_RELU)
    assert (out.cpu().double() - (torch.relu(ref) + R.double())).abs().max().item() <= tol * (scale + R.abs().max().item())
    sp = ops.gemm_tf32x3(As, W3, b.to(DEV), act=ops.ACT_RELU, split_out=True)
    assert sp.shape == (M, 2 * N)
    assert (sp[:, :N] + sp[:, N:]).cpu().double().sub(torch.relu(ref)).abs().max().item() <= tol * scale
    assert (sp[:, N:].abs() <= sp[:, :N].abs() * 2.0 ** -10 + 1e-30).all()      # lo is the tf32 remainder of hi
    # add[r % rows] prologue of the split
    add = torch.randn(37, K, generator=g)
    if M % 37 == 0 or True:
        m2 = (M // 37) * 37
        s2 = ops.split_rows_tf32(A[:m2].to(DEV).contiguous(), add=add.to(DEV))
        want = A[:m2] + add.repeat(m2 // 37, 1)
        assert torch.equal(s2[:, :K] + s2[:, K:], want.to(DEV))


def test_p2p_exchange_kernels_two_ranks_on_one_device():
    """Push / flag / acquire-merge kernels of the peer-memory key-split exchange with both 'ranks' on one GPU (their
    symmetric buffers are two local allocations; the peer-pointer tables are what a rendezvous would hand out): three
    forwards x two layers, results bit-equal to petr_attention_merge of the same partials, flags monotonic, no timeout."""
    from petr_b200._lib import call, ptr, stream_ptr
    B, Lq, Lk, world, L = 2, 100, 512, 2, 2
    qs, Kh, Vh, s = _flash_inputs(B, Lq, Lk, 2.0, 11)
    floats = ops.attention_partial_floats(B, Lq, 8)
    n = 2 * world * floats
    bufs = [torch.zeros(n + 64, dtype=torch.float32, device=DEV) for _ in range(world)]
    peer_bufs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    peer_flags = torch.tensor([b.data_ptr() + n * 4 for b in bufs], dtype=torch.int64, device=DEV)
    epochs = [torch.zeros(1, dtype=torch.int32, device=DEV) for _ in range(world)]
    counters = [torch.zeros(world, dtype=torch.int32, device=DEV) for _ in range(world)]
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    halves = ((0, 256), (256, Lk))
    sp = stream_ptr(DEV)
    for fwd in range(3):
        for r in range(world):
            call("petr_p2p_epoch_inc", ptr(epochs[r]), sp)
        for l in range(L):
            par = l & 1
            parts = []
            for r, (lo, hi) in enumerate(halves):
                slot = bufs[r][(par * world + r) * floats:(par * world + r + 1) * floats]
                ops.cross_attention_bf16(qs * (1.0 + 0.1 * l), Kh[:, :, lo:hi].contiguous(), Vh[:, :, lo:hi].contiguous(), None,
                                         B, Lq, hi - lo, 8, 1, partial=slot, write_partial=True)
                parts.append(slot.clone())
                call("petr_p2p_push_partial", ptr(slot), ptr(peer_bufs), ptr(peer_flags), world, r, floats, par,
                     ptr(epochs[r]), l, L, ptr(counters[r]), sp)
            want = ops.attention_merge(torch.cat(parts), world, B, Lq, 8)
            for r in range(world):
                got = torch.empty_like(want)
                flags = bufs[r][n:n + 2 * world].view(torch.int32)
                call("petr_attention_merge_p2p", ptr(bufs[r][par * world * floats:]), world, world, ptr(flags[par * world:]),
                     ptr(epochs[r]), l, L, ptr(err), ptr(got), got.stride(0), None, B, Lq, 8, sp)
                assert torch.equal(got, want), (fwd, l, r)
                assert flags[par * world:(par + 1) * world].tolist() == [(fwd + 1) * L + l + 1] * world
    assert int(err.item()) == 0
    assert all(int(c.sum()) == 0 for c in counters)            # the per-peer CTA counters wrapped back to zero


@pytest.mark.parametrize("B,Lk,layers", [(3, 33, 2), (2, 160, 1), (1, 257, 6), (5, 64, 3)])
def test_gemm_headsplit_tma_store_ragged(B, Lk, layers):
    """Head-split outputs (out_mode 2 and the hi / lo pair of out_mode 4) leave through swizzled slabs + TMA tensor stores
    where a warp's 32 rows are 32 consecutive keys of one sample, and through plain stores elsewhere (rows straddling a
    sample boundary, the ragged end of M): every key count / batch combination must give the same values."""
    g = torch.Generator().manual_seed(B * 1000 + Lk)
    M, K, N = B * Lk, 256, 256 * layers
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / 16).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    ref = (A.to(torch.bfloat16).double() @ W.to(torch.bfloat16).double().T + b.double()).float()
    refh = ref.view(B, Lk, layers, 8, 32).permute(2, 0, 3, 1, 4)                  # [layer][b][head][key][32]
    hs = ops.gemm_bf16_tc(A.to(torch.bfloat16), W.to(torch.bfloat16).contiguous(), b, headsplit_Lk=Lk)
    assert hs.shape == (layers, B, 8, Lk, 32)
    assert (hs.float() - refh).abs().max() <= 0.03 * max(1.0, ref.abs().max().item())
    # fp32-class pair: hi + lo against the exact product
    As, W3 = ops.split_rows(A), ops.x3_weight(W)
    ref32 = (A.double() @ W.double().T + b.double()).float().view(B, Lk, layers, 8, 32).permute(2, 0, 3, 1, 4)
    pair = ops.gemm_bf16_tc(As, W3, b, x3=True, headsplit_Lk=Lk, headsplit_pair=True)
    assert pair.shape == (2, layers, B, 8, Lk, 32)
    assert (pair[0].float() + pair[1].float() - ref32).abs().max() <= 6e-5 * max(1.0, ref32.abs().max().item())
    # and with the staging switched off (plain stores everywhere): identical bits
    _lib.set_tuning("gemm_staged", 0)
    try:
        hs0 = ops.gemm_bf16_tc(A.to(torch.bfloat16), W.to(torch.bfloat16).contiguous(), b, headsplit_Lk=Lk)
        pair0 = ops.gemm_bf16_tc(As, W3, b, x3=True, headsplit_Lk=Lk, headsplit_pair=True)
    finally:
        _lib.set_tuning("gemm_staged", 1)
    assert torch.equal(hs, hs0) and torch.equal(pair, pair0)


@pytest.mark.parametrize("M,K0,ksplit", [(900, 256, 3), (900, 2048, 8), (1800, 2048, 5), (130, 256, 2)])
def test_gemm_split_k_parts_sum_to_the_product(M, K0, ksplit):
    """Split-K (small-M query-side GEMMs): the parts written by petr_gemm_bf16_tc(ksplit) summed by
    petr_layernorm_multi_f32(nparts) give the LayerNorm of the unsplit product; deterministic."""
    g = torch.Generator().manual_seed(M + K0 + ksplit)
    A = torch.randn(M, K0, generator=g).to(DEV)
    W = (torch.randn(256, K0, generator=g) / math.sqrt(K0)).to(DEV)
    b, res = torch.randn(256, generator=g).to(DEV), torch.randn(M, 256, generator=g).to(DEV)
    gam, bet = torch.randn(256, generator=g).to(DEV), torch.randn(256, generator=g).to(DEV)
    As, W3 = ops.split_rows(A), ops.x3_weight(W)
    full = ops.gemm_bf16_tc(As, W3, b, residual=res, x3=True)
    parts = ops.gemm_bf16_tc(As, W3, b, residual=res, x3=True, ksplit=ksplit)
    assert parts.shape == (ksplit, M, 256)
    assert (parts.sum(0) - full).abs().max() <= 2e-5 * full.abs().max()
    Y0, Y1, Y2 = (torch.empty(M, 256, device=DEV) for _ in range(3))
    ops.layernorm_multi(full, gam, bet, Y=Y0)
    ops.layernorm_multi(parts, gam, bet, Y=Y1, nparts=ksplit)
    ops.layernorm_multi(ops.gemm_bf16_tc(As, W3, b, residual=res, x3=True, ksplit=ksplit), gam, bet, Y=Y2, nparts=ksplit)
    assert (Y1 - Y0).abs().max() <= 2e-5 * max(1.0, Y0.abs().max().item())
    assert torch.equal(Y1, Y2)
