"""Seeded synthetic head inputs: backbone features and nuScenes-like
``img_metas`` (lidar2img = K @ [R|t], float64 4x4 per camera) in the reference's own wire
format (lists of numpy arrays inside per-sample dicts, datasets/nuscenes_dataset.py:53-78).
Used by tests, smoke() and bench.py for BOTH arms so they see identical inputs.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch

from .configs import Workload


def _camera_extrinsic(i: int, rng: np.random.RandomState, ego_shift=None) -> np.ndarray:
    # lidar (x fwd, y left, z up) -> camera (x right, y down, z fwd), ring of 6 cameras
    swap = np.array([[0., -1., 0.], [0., 0., -1.], [1., 0., 0.]])
    a = -2.0 * math.pi * (i % 6) / 6.0
    rz = np.array([[math.cos(a), -math.sin(a), 0.], [math.sin(a), math.cos(a), 0.], [0., 0., 1.]])
    E = np.eye(4)
    E[:3, :3] = swap @ rz
    E[:3, 3] = rng.normal(0.0, 0.5, size=3)
    if ego_shift is not None:
        E[:3, 3] += E[:3, :3] @ ego_shift
    return E


def make_img_metas(wl: Workload, batch: int, seed: int = 0,
                   img_shape: Optional[Tuple[int, int]] = None,
                   pad_shape: Optional[Tuple[int, int]] = None) -> List[dict]:
    """Per-sample dicts with 'lidar2img', 'img_shape', 'pad_shape' (+ 'timestamp' for 12-view)."""
    rng = np.random.RandomState(1000 + seed)
    ih, iw = img_shape or (wl.img_h, wl.img_w)
    ph, pw = pad_shape or (ih, iw)
    K = np.array([[0.79 * iw, 0., iw / 2.0, 0.], [0., 0.79 * iw, ih / 2.0, 0.],
                  [0., 0., 1., 0.], [0., 0., 0., 1.]])
    metas = []
    for _ in range(batch):
        l2i = []
        ego = rng.normal(0.0, 1.0, size=3)
        for i in range(wl.num_cams):
            E = _camera_extrinsic(i, rng, ego_shift=ego if i >= 6 else None)
            l2i.append(K @ E)
        meta = dict(lidar2img=l2i,
                    img_shape=[(ih, iw, 3)] * wl.num_cams,
                    pad_shape=[(ph, pw, 3)] * wl.num_cams)
        if wl.num_cams > 6:
            half = wl.num_cams // 2
            meta['timestamp'] = [0.0] * half + [0.5] * (wl.num_cams - half)
        metas.append(meta)
    return metas


def make_features(wl: Workload, batch: int, seed: int = 0, scale: float = 1.0,
                  device='cpu', pin: bool = False) -> torch.Tensor:
    """x ~ N(0, scale^2), fp32 (B,N,Cin,H,W)."""
    g = torch.Generator(device='cpu').manual_seed(2000 + seed)
    x = torch.randn(batch, wl.num_cams, wl.in_channels, wl.feat_h, wl.feat_w, generator=g) * scale
    if pin:
        x = x.pin_memory()
    return x.to(device) if device != 'cpu' else x


def sharpen_state_dict(sd: dict, qk_scale: float = 8.0) -> dict:
    """Returns a copy whose cross-attention Wq/Wk rows are scaled so softmax is peaked
    (SURVEY.md section 4 vi): random-init attention is almost uniform and would not exercise
    online-softmax max tracking / rescaling."""
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        if k.endswith('attentions.1.attn.in_proj_weight'):
            E = out[k].shape[1]
            out[k][:2 * E] *= qk_scale
    return out


def random_state_dict(shapes: dict, seed: int = 0) -> dict:
    """Deterministic random weights for a head, generated key by key in sorted order from a
    seeded CPU generator (so the build container and the GPU box produce identical weights
    without shipping a 44 MB checkpoint).  Scales follow the reference's init_weights()
    (xavier-uniform matrices, reference_points ~ U(0,1), petr_head.py:276-284) but with non-zero
    biases and non-unit LayerNorm gains so every parameter matters in parity tests."""
    g = torch.Generator(device='cpu').manual_seed(3000 + seed)
    out = {}
    for k in sorted(shapes):
        if k.startswith('_'):
            continue
        shp = tuple(shapes[k])
        if k == 'code_weights':
            out[k] = torch.ones(shp)
        elif k == 'reference_points.weight':
            out[k] = torch.rand(shp, generator=g)
        elif len(shp) >= 2:
            fan_out, fan_in = shp[0], int(np.prod(shp[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            out[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
        elif 'norm' in k and k.endswith('weight') or (k.split('.')[-2] in ('1', '4') and
                                                      k.startswith('cls_branches') and k.endswith('weight')):
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            out[k] = 0.05 * torch.randn(shp, generator=g)
    # aliased (shared) branches of PETRHead must stay identical across the six copies
    for k in list(out):
        for br in ('cls_branches.', 'reg_branches.'):
            if k.startswith(br) and shapes.get('_shared_branches', False):
                rest = k.split('.', 2)[2]
                out[k] = out[f'{br}0.{rest}']
    return out


def head_state_dict(head, seed: int = 0) -> dict:
    """random_state_dict for a built head module (reference or petr_b200)."""
    shapes = {k: tuple(v.shape) for k, v in head.state_dict().items()}
    shared = head.cls_branches[0] is head.cls_branches[-1]
    sd = random_state_dict(dict(shapes, _shared_branches=shared) if shared else shapes, seed)
    return sd
