"""GPU parity of the training path: every backward kernel against torch autograd of the
same op, and the whole head -- forward in train mode, Hungarian matching, focal + L1 loss, gradients of EVERY parameter
and of the input features -- against autograd through the CPU oracle (oracle/petr_oracle.py head_forward + head_loss).

Tolerance (fp32 kernels, different summation order, atomics): kernel level max|g - g_ref| <= 1e-4..1e-3 of the tensor
scale; whole head: relative L2 <= 3e-3 per gradient tensor (see _close_l2)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import petr_b200 as pb
from oracle import petr_oracle as po
from petr_b200 import ops, synth, train
from petr_b200.ops import ACT_NONE, ACT_RELU
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, rel=2e-3, name=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = b.abs().max().item()
    err = (a - b).abs().max().item()
    assert math.isfinite(err) and err <= rel * scale + 1e-6, (name, err, scale)
    return err / max(scale, 1e-30)


def _close_l2(a, b, name="", floor=1e-6, l2_bound=3e-3, max_bound=2e-2):
    """Whole-head gradients: a ReLU unit whose fp32 pre-activation sits within rounding of 0 may switch side between two
    correct implementations, which moves ONE row of a weight gradient by a finite amount.  Bound: relative L2 error
    <= 3e-3 and max error <= 2e-2 of the tensor's scale (a wrong kernel is off by O(1) in both).  `floor`: absolute
    noise level (1e-7 of the largest gradient of the model) below which a tensor is numerically zero -- e.g. the Q/K
    projections of the first self-attention, whose values are all zero.  With the Linear layers on the bf16x3 tensor-core
    tier (operands carry 16 mantissa bits: ~1e-5 of the output scale instead of 1e-7) about a hundred times more units
    sit within rounding of 0, and the cross-attention concentrates the key-side gradient on few tokens, so one
    switched unit can carry a few percent of a bias gradient: relative L2 <= 1e-2, max <= 5e-2 of the scale there."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = b.abs().max().item()
    err = (a - b).abs().max().item()
    assert math.isfinite(err) and err <= max_bound * scale + floor, (name, err, scale)
    dn, bn = (a - b).norm().item(), b.norm().item()
    l2 = dn / max(bn, 1e-30)
    assert l2 <= l2_bound or dn <= floor * math.sqrt(a.numel()), (name, err, scale, l2)
    return l2 if scale > 100 * floor else 0.0


def _leaf(t):
    return t.clone().to(DEV).requires_grad_(True)


@pytest.mark.parametrize("M,K,N,rows,tc", [(333, 192, 100, 111, False), (640, 192, 128, 128, True)])
def test_linear_backward_variants(M, K, N, rows, tc):
    """Ragged shape: FFMA kernels; 64-aligned shape: bf16x3 tcgen05 GEMMs for the forward and the input gradient."""
    assert train._tc_ok(M, N, K) == tc and train._tc_ok(M, K, N) == tc
    g = torch.Generator().manual_seed(0)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    A2, R = torch.randn(rows, K, generator=g), torch.randn(M, N, generator=g)
    dY = torch.randn(M, N, generator=g)
    for act, use_a2, use_res in ((ACT_NONE, True, True), (ACT_RELU, True, False), (ACT_NONE, False, False)):
        mine = [_leaf(t) for t in (A, W, b, A2, R)]
        out = train.linear(mine[0], mine[1], mine[2], A2=mine[3] if use_a2 else None,
                           residual=mine[4] if use_res else None, act=act)
        ref_in = [t.clone().double().requires_grad_(True) for t in (A, W, b, A2, R)]
        a = ref_in[0] + (ref_in[3].repeat(M // rows, 1) if use_a2 else 0)
        y = F.linear(a, ref_in[1], ref_in[2])
        if act == ACT_RELU:
            # the reference takes the kernel's own ReLU mask: a pre-activation within rounding of 0 may sit on either
            # side in two correct implementations, and one switched unit moves a row of dA by |dY| |W|
            assert (F.relu(y) - out.detach().cpu().double()).abs().max() <= 3e-5 * y.abs().max()
            y = y * (out.detach().cpu() > 0)
        if use_res:
            y = y + ref_in[4]
        y.backward(dY.double())
        _close(out, y, 3e-5 if tc else 1e-5, "linear fwd")     # bf16x3: operands carry 16 mantissa bits
        out.backward(dY.to(DEV))
        for i, nm in enumerate(("dA", "dW", "db", "dA2", "dres")):
            if (nm == "dA2" and not use_a2) or (nm == "dres" and not use_res):
                continue
            _close(mine[i].grad, ref_in[i].grad, 1e-4, nm)
    # conv1x1 over an NCHW map (input_proj): gradients w.r.t. the map and the conv weight
    x = torch.randn(3, 64, 5, 7, generator=g)
    Wc, bc = torch.randn(32, 64, 1, 1, generator=g) * 0.1, torch.randn(32, generator=g)
    xr, Wr, br = (t.clone().double().requires_grad_(True) for t in (x, Wc, bc))
    yr = F.conv2d(xr, Wr, br).permute(0, 2, 3, 1).reshape(-1, 32)
    d = torch.randn(yr.shape, generator=g)
    yr.backward(d.double())
    xm, Wm, bm = _leaf(x), _leaf(Wc), _leaf(bc)
    ym = train.linear(xm, Wm, bm, nchw_hw=35)
    _close(ym, yr, 1e-5)
    ym.backward(d.to(DEV))
    _close(xm.grad, xr.grad, 1e-4, "dx nchw")
    _close(Wm.grad, Wr.grad, 1e-4, "dWconv")
    _close(bm.grad, br.grad, 1e-4, "dbconv")


@pytest.mark.parametrize("relu,n2n", [(False, False), (True, False), (False, True)])
def test_layernorm_backward(relu, n2n):
    g = torch.Generator().manual_seed(1)
    X = torch.randn(77, 256, generator=g) * 2 + 0.3
    if n2n:
        X[5] = float("nan")
    gam, bet, dY = torch.randn(256, generator=g), torch.randn(256, generator=g), torch.randn(77, 256, generator=g)
    ok = ~torch.isnan(X).any(1)
    xr, gr, br = X[ok].clone().double().requires_grad_(True), gam.clone().double().requires_grad_(True), \
        bet.clone().double().requires_grad_(True)
    y = F.layer_norm(xr, (256,), gr, br, 1e-5)
    if relu:
        y = F.relu(y)
    y.backward(dY[ok].double())
    xm, gm, bm = _leaf(X), _leaf(gam), _leaf(bet)
    ym = train.layernorm(xm, gm, bm, 1e-5, relu, n2n)
    _close(ym[ok.to(DEV)], y, 1e-5)
    ym.backward(dY.to(DEV))
    _close(xm.grad[ok.to(DEV)], xr.grad, 1e-4, "dx")
    _close(gm.grad, gr.grad, 1e-4, "dgamma")
    _close(bm.grad, br.grad, 1e-4, "dbeta")
    if n2n:
        assert (xm.grad[5] == 0).all()


@pytest.mark.parametrize("B,Lq,Lk,masked", [(2, 70, 333, True), (1, 130, 130, False), (3, 64, 1500, True)])
def test_attention_backward(B, Lq, Lk, masked):
    g = torch.Generator().manual_seed(2)
    H, E = 8, 256
    Q, K, V = (torch.randn(B * L, E, generator=g) for L in (Lq, Lk, Lk))
    km = (torch.rand(B, Lk, generator=g) < 0.3) if masked else None
    dO = torch.randn(B * Lq, E, generator=g)
    qr, kr, vr = (t.clone().double().requires_grad_(True) for t in (Q, K, V))

    def heads(t, L):
        return t.view(B, L, H, 32).transpose(1, 2)
    s = heads(qr, Lq) @ heads(kr, Lk).transpose(-1, -2) / math.sqrt(32)
    if km is not None:
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    o = (s.softmax(-1) @ heads(vr, Lk)).transpose(1, 2).reshape(B * Lq, E)
    o.backward(dO.double())
    qm, kmm, vm = _leaf(Q), _leaf(K), _leaf(V)
    om = train._Attention.apply(qm, kmm, vm, km.to(torch.uint8).to(DEV) if km is not None else None, B, Lq, Lk, H, 0.0, 0)
    _close(om, o, 1e-5)
    om.backward(dO.to(DEV))
    _close(qm.grad, qr.grad, 2e-4, "dQ")
    _close(kmm.grad, kr.grad, 2e-4, "dK")
    _close(vm.grad, vr.grad, 2e-4, "dV")


def test_attention_dropout_consistency():
    """With dropout the forward and both backward kernels must see the SAME mask: O is linear in V, so the directional
    derivative <dO, O(V + eps dV')> - <dO, O(V)> must equal eps <dV, dV'> exactly up to fp32 rounding."""
    g = torch.Generator().manual_seed(3)
    B, Lq, Lk, H, E = 2, 65, 200, 8, 256
    Q, K, V, dVp = (torch.randn(B * L, E, generator=g).to(DEV) for L in (Lq, Lk, Lk, Lk))
    dO = torch.randn(B * Lq, E, generator=g).to(DEV)
    p, seed = 0.3, 1234
    vm = V.clone().requires_grad_(True)
    qm = Q.clone().requires_grad_(True)
    o = train._Attention.apply(qm, K, vm, None, B, Lq, Lk, H, p, seed)
    o.backward(dO)
    o2 = train._Attention.apply(Q, K, V + dVp, None, B, Lq, Lk, H, p, seed)
    lhs = ((o2 - o.detach()) * dO).sum().item()
    rhs = (vm.grad * dVp).sum().item()
    assert abs(lhs - rhs) <= 2e-3 * abs(rhs) + 1e-3, (lhs, rhs)
    # the mask really drops ~p of the probability mass and rescales the rest: E[O] is unchanged
    o0 = train._Attention.apply(Q, K, V, None, B, Lq, Lk, H, 0.0, 0)
    outs = torch.stack([train._Attention.apply(Q, K, V, None, B, Lq, Lk, H, p, 100 + i) for i in range(64)]).mean(0)
    assert (outs - o0).abs().mean() <= 0.05 * o0.abs().mean() + 0.02
    assert not torch.equal(o.detach(), o0)
    # dQ under dropout against torch with the mask recovered from a V = identity probe is covered by the head test;
    # here: finite and non-zero
    assert torch.isfinite(qm.grad).all() and qm.grad.abs().max() > 0


def test_dropout_and_small_ops_backward():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(50, 256, generator=g).to(DEV)
    r = torch.randn(50, 256, generator=g).to(DEV)
    xm, rm = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    train.manual_seed(7)
    y = train.dropout(xm, 0.25, residual=rm)
    keep = ((y - r) != 0)
    frac = keep.float().mean().item()
    assert 0.70 <= frac <= 0.80, frac
    assert torch.allclose((y - r)[keep], (x / 0.75)[keep], rtol=1e-6, atol=1e-6)
    y.backward(torch.ones_like(y))
    assert torch.equal(xm.grad != 0, keep) and torch.allclose(xm.grad[keep], torch.full_like(xm.grad[keep], 1 / 0.75))
    assert torch.equal(rm.grad, torch.ones_like(r))
    # pos2posemb3d
    ref = torch.rand(40, 3, generator=g)
    dim_t = 10000 ** (2 * (torch.arange(128, dtype=torch.float32) // 2) / 128)
    rr = ref.clone().double().requires_grad_(True)
    e = po.pos2posemb3d(rr.float()).double() if False else None
    pr = rr * (2 * math.pi)
    parts = []
    for ax in (1, 0, 2):
        a = pr[:, ax, None] / dim_t.double()
        parts.append(torch.stack((a[:, 0::2].sin(), a[:, 1::2].cos()), -1).flatten(-2))
    e = torch.cat(parts, -1)
    d = torch.randn(e.shape, generator=g)
    e.backward(d.double())
    rm_ = _leaf(ref)
    em = train._Pos2PosEmb.apply(rm_, dim_t.to(DEV))
    _close(em, e, 1e-4)
    em.backward(d.to(DEV))
    _close(rm_.grad, rr.grad, 1e-3, "dref")
    # decode boxes (with a PETRv2 time divisor)
    L, B, Q, cs = 2, 3, 40, 10
    reg = torch.randn(L * B * Q, cs, generator=g)
    td = torch.tensor([0.5, 0.4, 0.6])
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    regr, refr = reg.clone().double().requires_grad_(True), ref.clone().double().requires_grad_(True)
    rf = po.inverse_sigmoid(refr)
    t = regr.view(L, B, Q, cs)
    xy = (t[..., 0:2] + rf[:, 0:2]).sigmoid()
    z = (t[..., 4:5] + rf[:, 2:3]).sigmoid()
    vel = t[..., 8:] / td.double()[None, :, None, None]
    outr = torch.cat([xy[..., 0:1] * (pc[3] - pc[0]) + pc[0], xy[..., 1:2] * (pc[4] - pc[1]) + pc[1], t[..., 2:4],
                      z * (pc[5] - pc[2]) + pc[2], t[..., 5:8], vel], -1).view(-1, cs)
    d = torch.randn(outr.shape, generator=g)
    outr.backward(d.double())
    regm, refm = _leaf(reg), _leaf(ref)
    outm = train._DecodeBoxes.apply(regm, refm, pc, td.to(DEV), L, B, Q)
    _close(outm, outr, 1e-5)
    outm.backward(d.to(DEV))
    _close(regm.grad, regr.grad, 1e-4, "dreg")
    _close(refm.grad, refr.grad, 1e-3, "dref decode")


def test_norm_scale_and_learned_pos_backward():
    """NormedLinear operand scaling (both group layouts, a dead channel included) and the LearnedPositionalEncoding3D
    table gradients against torch autograd of the published formulas."""
    g = torch.Generator().manual_seed(3)
    B, Q, E = 2, 37, 70
    for power in (1.0, 2.0):
        x = torch.randn(B, Q, E, generator=g)
        x[0, :, 5] = 0.0                                   # a channel ReLU left dead for every query
        dy = torch.randn(B, Q, E, generator=g)
        for along in (True, False):
            xr = x.clone().requires_grad_(True)
            dim = 1 if along else 2
            yr = xr / (xr.norm(dim=dim, keepdim=True).pow(power) + 1e-6) * 20.0
            yr.backward(dy)
            xm = _leaf(x.view(B * Q, E))
            ym = train._NormScale.apply(xm, B, Q, along, power, 1e-6, 20.0)
            _close(ym, yr.view(B * Q, E), 1e-5, "norm_scale fwd")
            ym.backward(dy.view(B * Q, E).to(DEV))
            live = torch.ones(B, Q, E, dtype=torch.bool)
            if along:
                live[0, :, 5] = False                      # zero-norm group: dOut * mult / eps on both sides
                _close(xm.grad.view(B, Q, E)[0, :, 5], xr.grad[0, :, 5], 1e-5, "dead channel")
            _close(xm.grad.view(B, Q, E)[live], xr.grad[live], 1e-4, f"norm_scale bwd along={along} p={power}")
    G, H, W, Fn = 3, 5, 7, 48
    row, col = torch.randn(9, Fn, generator=g), torch.randn(11, Fn, generator=g)
    d = torch.randn(G * H * W, 2 * Fn, generator=g)
    rr, cr = row.clone().requires_grad_(True), col.clone().requires_grad_(True)
    ref = torch.cat((cr[:W].unsqueeze(0).repeat(H, 1, 1), rr[:H].unsqueeze(1).repeat(1, W, 1)), -1)
    ref = ref.unsqueeze(0).repeat(G, 1, 1, 1).reshape(G * H * W, 2 * Fn)
    ref.backward(d)
    rm, cm = _leaf(row), _leaf(col)
    out = train._LearnedPos.apply(rm, cm, G, H, W)
    assert torch.equal(out.cpu(), ref.detach())
    out.backward(d.to(DEV))
    _close(rm.grad, rr.grad, 1e-5, "d row_embed")
    _close(cm.grad, cr.grad, 1e-5, "d col_embed")


def _gts(B, n, seed):
    g = torch.Generator().manual_seed(seed)
    out, lab = [], []
    for b in range(B):
        k = n if b != 1 else max(n - 2, 0)
        ctr = torch.rand(k, 3, generator=g) * torch.tensor([80.0, 80.0, 6.0]) - torch.tensor([40.0, 40.0, 4.0])
        dims = torch.rand(k, 3, generator=g) * 3 + 0.5
        rot = (torch.rand(k, 1, generator=g) - 0.5) * 6
        vel = torch.randn(k, 2, generator=g)
        out.append(torch.cat([ctr, dims, rot, vel], 1))
        lab.append(torch.randint(0, 10, (k,), generator=g))
    return out, lab


def test_loss_kernels_and_cost_matrix():
    g = torch.Generator().manual_seed(5)
    Q, C, G = 64, 10, 7
    cls, box = torch.randn(Q, C, generator=g) * 2, torch.randn(Q, 10, generator=g)
    gts, lab = _gts(1, G, 9)
    ref = po.hungarian_cost(cls, box, gts[0], lab[0])
    ng = torch.empty((G, 10), device=DEV)
    gd, cd, bd, ld = gts[0].to(DEV), cls.to(DEV), box.to(DEV), lab[0].to(DEV)       # keep the device copies alive
    train.call("petr_normalize_bbox", train.ptr(gd), G, 9, train.ptr(ng), train.stream_ptr(DEV))
    _close(ng, po.normalize_bbox(gts[0]), 1e-6)
    cost = torch.empty((Q, G), device=DEV)
    train.call("petr_hungarian_cost", train.ptr(cd), train.ptr(bd), train.ptr(ld), train.ptr(ng),
               Q, G, C, 10, 10, 2.0, 0.25, 0.25, 2.0, 1e-12, train.ptr(cost), train.stream_ptr(DEV))
    _close(cost, ref, 1e-5, "cost")
    # focal loss + gradient, background rows included
    labels = torch.randint(0, C + 1, (Q,), generator=g)
    xr = cls.clone().double().requires_grad_(True)
    lr = po.sigmoid_focal_loss(xr, labels, C).sum() * 0.37
    lr.backward()
    xm = _leaf(cls)
    lm = train._FocalLoss.apply(xm, labels.to(DEV), None, 0.25, 2.0, 0.37)
    _close(lm, lr, 1e-5)
    (lm * 2.0).backward()
    _close(xm.grad, xr.grad * 2.0, 1e-4, "dfocal")
    # L1 with non-finite target rows skipped
    T = torch.randn(Q, 10, generator=g)
    T[3, 2] = float("-inf")
    Wt = torch.rand(Q, 10, generator=g)
    ok = torch.isfinite(T).all(-1)
    br = box.clone().double().requires_grad_(True)
    l1r = ((br[ok] - T[ok].double()).abs() * Wt[ok].double()).sum() * 0.11
    l1r.backward()
    bm = _leaf(box)
    l1m = train._L1Loss.apply(bm, T.to(DEV), Wt.to(DEV), 0.11)
    _close(l1m, l1r, 1e-5)
    l1m.backward()
    _close(bm.grad, br.grad, 1e-5, "dl1")


TRAIN_CASES = {
    "tiny": dict(workload="tiny", batch=2, num_query=64, seed=0),
    "tiny_padded": dict(workload="tiny", batch=3, num_query=64, seed=1, img_shape=(40, 130), pad_shape=(64, 160)),
    "small": dict(workload="small", batch=2, num_query=64, seed=5),
    "tiny_v2": dict(workload="tiny-v2", batch=1, num_query=64, seed=4),
}


@pytest.mark.parametrize("tc_gemm", [False, True], ids=["ffma", "bf16x3"])
@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_head_loss_and_gradients_match_oracle_autograd(name, tc_gemm):
    """Both GEMM tiers of the training path: exact FFMA (train.TC_GEMM = False; pins the gradients) and the default
    bf16x3 tcgen05 GEMMs for the large Linear layers."""
    _check_head_gradients(name, TRAIN_CASES[name], tc_gemm=tc_gemm)


# the reference branches no shipped config takes (petr_head.py:231-234, 396-420), in train mode
TRAIN_VARIANTS = {
    "learned_pos": (dict(with_multiview=False, positional_encoding=dict(
        type='LearnedPositionalEncoding3D', num_feats=128, row_num_embed=50, col_num_embed=50)),
        dict(with_multiview=False, pos_kind="learned")),
    "sine2d_pos": (dict(with_multiview=False, positional_encoding=dict(
        type='SinePositionalEncoding', num_feats=128, normalize=True)), dict(with_multiview=False, pos_kind="sine2d")),
    "sine3d_only": (dict(with_position=False), dict(with_position=False)),
    "normedlinear": (dict(normedlinear=True), dict(normedlinear=True)),
}


@pytest.mark.parametrize("variant", list(TRAIN_VARIANTS))
def test_head_gradients_config_branches(variant):
    hc, oc = TRAIN_VARIANTS[variant]
    _check_head_gradients(variant, TRAIN_CASES["tiny"], hc, oc)


def _check_head_gradients(name, case, head_over=None, oracle_over=None, tc_gemm=True):
    saved = train.TC_GEMM
    train.TC_GEMM = tc_gemm
    try:
        _check_head_gradients_impl(name, case, head_over, oracle_over, 1e-2 if tc_gemm else 3e-3)
    finally:
        train.TC_GEMM = saved


def _check_head_gradients_impl(name, case, head_over, oracle_over, l2_bound):
    wl, feats, metas = cases.inputs(case)
    cfg = cases.head_cfg(case)
    ocfg = cases.oracle_cfg(case)
    if head_over:
        cfg.update(head_over)
        ocfg.update(oracle_over)
        head = pb.build_head(cfg)
        sd = synth.head_state_dict(head, case["seed"])
        if cfg.get("normedlinear"):    # temperature 20 on unit-norm operands: keep the logits in a sane range
            for k in sd:
                if k.startswith("cls_branches.") and k.endswith(".6.weight"):
                    sd[k] = sd[k] * 0.05
    else:
        head = pb.build_head(cfg)
        sd = cases.weights(case, head)
    head.load_state_dict(sd, strict=True)
    head = head.to(DEV)
    B = feats.shape[0]
    gts, lab = _gts(B, 6, 17)
    # oracle: autograd through the CPU restatement
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    xr = feats.clone().requires_grad_(True)
    outr = po.head_forward(sdr, xr, metas, ocfg)
    lossr, mr = po.head_loss(outr, gts, lab, code_weights=sd["code_weights"].tolist())
    sum(lossr.values()).backward()
    # ours: differentiable forward without dropout (== eval function), loss, backward
    xm = feats.clone().to(DEV).requires_grad_(True)
    ops.reset_launch_count()
    outm = head.forward_autograd([xm], metas, dropout=False)
    for k in ("all_cls_scores", "all_bbox_preds"):
        assert (outm[k].detach().cpu() - outr[k].detach()).abs().max() <= 2e-3 * max(1.0, float(outr[k].abs().max()))
    lossm = head.loss([t.to(DEV) for t in gts], [t.to(DEV) for t in lab], outm)
    assert list(lossm) == list(lossr)
    mm = head._loss_impl.assign(outm["all_cls_scores"], outm["all_bbox_preds"], [t.to(DEV) for t in gts],
                                [t.to(DEV) for t in lab])
    for l in range(len(mr)):
        for b in range(B):
            assert np.array_equal(mm[l][b][0], mr[l][b][0]) and np.array_equal(mm[l][b][1], mr[l][b][1]), (l, b)
    for k in lossr:
        _close(lossm[k], lossr[k], 1e-4, k)
    sum(lossm.values()).backward()
    assert ops.launch_count() > 150          # forward + loss launches (the counter is per thread; autograd runs backward on its own)
    floor = 1e-7 * max(float(v.grad.abs().max()) for v in sdr.values() if torch.is_tensor(v) and v.grad is not None)
    max_bound = 5e-2 if l2_bound > 3e-3 else 2e-2
    _close_l2(xm.grad, xr.grad, "d features", floor, l2_bound, max_bound)
    worst = 0.0
    named = dict(head.named_parameters())
    checked = 0
    for k, p in named.items():
        if not p.requires_grad:
            continue
        gr = sdr[k].grad
        if gr is None:
            assert p.grad is None or p.grad.abs().max() == 0, k
            continue
        # PETRHead shares one branch module between the six prediction layers: the oracle's state_dict holds six
        # aliased copies, each with its own gradient -- the module's gradient is their sum
        if k.startswith(("cls_branches.0.", "reg_branches.0.")) and head.cls_branches[0] is head.cls_branches[-1]:
            gr = sum(sdr[k.replace(".0.", f".{l}.", 1)].grad for l in range(6))
        assert p.grad is not None, k
        worst = max(worst, _close_l2(p.grad, gr, k, floor, l2_bound, max_bound))
        checked += 1
    print(f"{name}: {checked} parameter gradients checked, worst relative L2 error {worst:.2e}")
    assert checked >= 100


def test_train_mode_step_reduces_loss():
    """.train() forward (dropout 0.1 active as configured), loss, backward, SGD: finite everywhere and the loss falls."""
    case = TRAIN_CASES["tiny"]
    wl, feats, metas = cases.inputs(case)
    head = pb.build_head(cases.head_cfg(case))
    head.load_state_dict(cases.weights(case, head), strict=True)
    head = head.to(DEV).train()
    gts, lab = _gts(feats.shape[0], 5, 3)
    gts, lab = [t.to(DEV) for t in gts], [t.to(DEV) for t in lab]
    opt = torch.optim.SGD([p for p in head.parameters() if p.requires_grad], lr=2e-4)
    x = feats.to(DEV)
    train.manual_seed(11)
    hist = []
    for it in range(8):
        opt.zero_grad()
        out = head([x], metas)
        assert out["all_cls_scores"].requires_grad
        losses = head.loss(gts, lab, out)
        tot = sum(losses.values())
        tot.backward()
        assert all(torch.isfinite(p.grad).all() for p in head.parameters() if p.grad is not None)
        opt.step()
        hist.append(float(tot))
    print("loss history", [round(h, 3) for h in hist])
    assert all(math.isfinite(h) for h in hist) and min(hist[-3:]) < hist[0]
    # back to inference: the engine sees the updated parameters (version counters) and runs the eval path
    head.eval()
    o = head([x], metas)
    assert not o["all_cls_scores"].requires_grad and torch.isfinite(o["all_cls_scores"]).all()


def test_loss_with_a_sample_without_ground_truth():
    """hungarian_assigner_3d.py:100-106: a sample with no boxes assigns every query to background; only the focal loss
    sees it.  Losses and the feature gradient must still match the oracle."""
    case = TRAIN_CASES["tiny"]
    wl, feats, metas = cases.inputs(case)
    head = pb.build_head(cases.head_cfg(case))
    sd = cases.weights(case, head)
    head.load_state_dict(sd, strict=True)
    head = head.to(DEV)
    gts, lab = _gts(feats.shape[0], 2, 23)            # sample 1: 2 - 2 = 0 boxes
    assert gts[1].shape[0] == 0 and gts[0].shape[0] == 2
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    xr = feats.clone().requires_grad_(True)
    lossr, _ = po.head_loss(po.head_forward(sdr, xr, metas, cases.oracle_cfg(case)), gts, lab,
                            code_weights=sd["code_weights"].tolist())
    sum(lossr.values()).backward()
    xm = feats.clone().to(DEV).requires_grad_(True)
    lossm = head.loss([t.to(DEV) for t in gts], [t.to(DEV) for t in lab], head.forward_autograd([xm], metas, dropout=False))
    for k in lossr:
        _close(lossm[k], lossr[k], 1e-4, k)
    sum(lossm.values()).backward()
    _close_l2(xm.grad, xr.grad, "d features", 1e-7 * float(xr.grad.abs().max()))
