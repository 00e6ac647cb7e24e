"""Upstream neck tail of the head hot path: `CPFPN`
(projects/mmdet3d_plugin/models/necks/cp_fpn.py:17-210) -- 1x1 lateral convs, nearest top-down adds and the single
3x3 conv the reference keeps on level 0 -- with the reference's constructor kwargs and state_dict layout
(`lateral_convs.{i}.conv.*`, `fpn_convs.0.conv.*`).

All maps are kept token-major ((g, y, x) rows, channels contiguous) between the kernels: a lateral conv is one GEMM
straight from the NCHW input, the 3x3 conv is an im2col pass + one GEMM with the weight repacked to [Cout][tap][Cin].
'bf16' precision runs the GEMMs on tcgen05 (bf16 operands, fp32 accumulate), 'fp32' on the exact FFMA kernel.
CUDA tensors only (no CPU / PyTorch fallback).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import ops
from .registry import NECKS


class ConvModule(nn.Module):
    """mmcv ConvModule without norm / activation (every PETR config: norm_cfg=None, act_cfg=None): the parameter
    container that yields the `<name>.conv.weight` / `.conv.bias` checkpoint keys."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding)


@NECKS.register_module()
class CPFPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest'),
                 init_cfg=dict(type='Xavier', layer='Conv2d', distribution='uniform')):
        super().__init__()
        assert isinstance(in_channels, list)
        if conv_cfg is not None or norm_cfg is not None or act_cfg is not None:
            raise NotImplementedError("petr_b200 CPFPN: plain convs only (conv_cfg / norm_cfg / act_cfg = None, "
                                      "as in every PETR config)")
        if upsample_cfg.get('mode', 'nearest') != 'nearest' or 'scale_factor' in upsample_cfg:
            raise NotImplementedError("petr_b200 CPFPN: upsample_cfg=dict(mode='nearest') only")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        self.relu_before_extra_convs = relu_before_extra_convs
        self.no_norm_on_lateral = no_norm_on_lateral
        self.fp16_enabled = False
        self.upsample_cfg = dict(upsample_cfg)
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels)
            assert num_outs == end_level - start_level
        self.start_level = start_level
        self.end_level = end_level
        self.add_extra_convs = add_extra_convs
        if num_outs > self.backbone_end_level - start_level:
            raise NotImplementedError("petr_b200 CPFPN: extra output levels (max-pool / extra convs) are not used by "
                                      "any PETR config and not built")
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1))
            if i == 0:
                self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1))
        self.precision = 'bf16'
        self._cache = None

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    # ---- packed weights (rebuilt when a parameter changes) ----
    def _packed(self, device):
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (str(device), self.precision)
        if self._cache is not None and self._cache[0] == sig:
            return self._cache[1]
        bf = self.precision == 'bf16'
        lat = []
        for m in self.lateral_convs:
            w = m.conv.weight.detach().to(device=device, dtype=torch.float32).flatten(1).contiguous()
            lat.append((w.to(torch.bfloat16).contiguous() if bf else w,
                        m.conv.bias.detach().to(device=device, dtype=torch.float32).contiguous()))
        fpn = None
        if len(self.fpn_convs):
            c = self.fpn_convs[0].conv
            # [Cout][Cin][ky][kx] -> [Cout][tap = 3 ky + kx][Cin]: the column order petr_im2col3x3_tokens writes
            w = c.weight.detach().to(device=device, dtype=torch.float32).permute(0, 2, 3, 1).reshape(
                c.weight.shape[0], -1).contiguous()
            fpn = (w.to(torch.bfloat16).contiguous() if bf else w,
                   c.bias.detach().to(device=device, dtype=torch.float32).contiguous())
        self._cache = (sig, dict(lat=lat, fpn=fpn))
        return self._cache[1]

    @torch.no_grad()
    def forward_tokens(self, inputs: Sequence[torch.Tensor]):
        """cp_fpn.py:159-210 on token-major maps.  Returns [(tokens fp32 (G*h*w, Cout), G, h, w)] per output level."""
        assert len(inputs) == len(self.in_channels)
        if not all(x.is_cuda for x in inputs):
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        P = self._packed(inputs[0].device)
        bf = self.precision == 'bf16'
        lats = []
        for i, (w, b) in enumerate(P["lat"]):
            x = inputs[i + self.start_level].float().contiguous()
            G, C, h, wd = x.shape
            if bf:
                if C % 64 != 0:
                    raise NotImplementedError("tcgen05 path needs in_channels % 64 == 0 (use precision='fp32')")
                t = ops.gemm_bf16_tc(ops.nchw_to_tokens_bf16(x), w, b, name="gemm_fpn_lateral")
            else:
                t = ops.gemm_f32(x, w, b, nchw_hw=h * wd)
            lats.append((t, G, h, wd))
        for i in range(len(lats) - 1, 0, -1):
            (dst, G, H, W), (src, _, h, wd) = lats[i - 1], lats[i]
            ops.upsample_add_tokens(dst, src, G, H, W, h, wd)
        outs = []
        for i, (t, G, h, wd) in enumerate(lats):
            if i == 0 and self.start_level == 0 and P["fpn"] is not None:
                w, b = P["fpn"]
                if bf:
                    t = ops.gemm_bf16_tc(ops.im2col3x3_tokens(t, G, h, wd, "bf16"), w, b, name="gemm_fpn_conv3x3")
                else:
                    t = ops.gemm_f32(ops.im2col3x3_tokens(t, G, h, wd, "f32"), w, b)
            outs.append((t, G, h, wd))
        return outs

    @torch.no_grad()
    def forward(self, inputs):
        """Same contract as the reference: tuple of NCHW fp32 maps (G, Cout, h, w)."""
        return tuple(ops.tokens_to_nchw(t, G, h * w).view(G, self.out_channels, h, w)
                     for t, G, h, w in self.forward_tokens(inputs))
