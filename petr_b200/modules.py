"""Host-side mirror of the reference's plugin interface for the head hot path.

Same registered names, constructor kwargs, forward signatures and ``state_dict`` layout as
  PETRHead                     projects/mmdet3d_plugin/models/dense_heads/petr_head.py:47-468
  PETRv2Head                   .../dense_heads/petrv2_head.py:99-540
  PETRTransformer              .../models/utils/petr_transformer.py:34-109
  PETRTransformerDecoder       .../petr_transformer.py:401-447
  PETRTransformerDecoderLayer  .../petr_transformer.py:113-224   (mmcv BaseTransformerLayer)
  PETRMultiheadAttention       .../petr_transformer.py:228-367
  MultiheadAttention / FFN     mmcv-full 1.4.0 bricks of the same names
  SinePositionalEncoding3D     .../models/utils/positional_encoding.py:15-110
  NMSFreeCoder                 .../core/bbox/coders/nms_free_coder.py:18-119 (pc_range holder + decode)
so a reference checkpoint loads with ``strict=True`` and any projects/configs/petr* or petrv2*
``pts_bbox_head`` dict builds through ``registry.build_head``.

The modules own parameters only; every forward launches kernels of libpetr_b200.so (inference
forward; there is no CPU or PyTorch fallback -- CPU tensors raise).
"""
from __future__ import annotations

import copy
import math
import warnings
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .engine import HeadEngine, _sine_args
from .ops import ACT_RELU
from .registry import (ATTENTION, BBOX_CODERS, FEEDFORWARD_NETWORK, HEADS, POSITIONAL_ENCODING,
                       TRANSFORMER, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE,
                       build_attention, build_bbox_coder, build_feedforward_network,
                       build_positional_encoding, build_transformer,
                       build_transformer_layer, build_transformer_layer_sequence)


def _require_eval(m: nn.Module):
    if m.training:
        raise RuntimeError(f"{type(m).__name__}: petr_b200 implements the inference forward only; "
                           "call .eval() (training/backward is listed as a later row in DESIGN.md)")


def _rows(t: torch.Tensor, batch_first: bool):
    """(L,B,E) seq-first (or (B,L,E)) -> contiguous (B*L, E) rows, plus (B, L)."""
    if not batch_first:
        t = t.transpose(0, 1)
    B, L, E = t.shape
    return t.contiguous().view(B * L, E), B, L


def set_precision(module: nn.Module, precision: str) -> nn.Module:
    """Sets the precision path ('bf16' = tcgen05 kernels, 'fp32' = exact FFMA kernels) of every attention / FFN brick,
    neck and head below `module`."""
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    for m in module.modules():
        if hasattr(type(m), "precision") or "precision" in m.__dict__:
            m.precision = precision
    return module


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()
        self._is_init = True


# ------------------------------------------------------------------------------------------
# attention / FFN bricks
# ------------------------------------------------------------------------------------------
@ATTENTION.register_module()
class PETRMultiheadAttention(BaseModule):
    """petr_transformer.py:228-367: pos-add, packed in-proj, softmax attention, out-proj,
    identity residual.  ``self.attn`` is an ``nn.MultiheadAttention`` used as the parameter
    container so checkpoint keys match (``attn.in_proj_weight`` ...)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        dropout_layer = dict(dropout_layer) if dropout_layer else dropout_layer
        if 'dropout' in kwargs:
            warnings.warn('The arguments `dropout` in MultiheadAttention has been deprecated, now '
                          'you can separately set `attn_drop`(float), proj_drop(float), and '
                          '`dropout_layer`(dict) ', DeprecationWarning)
            attn_drop = kwargs['dropout']
            dropout_layer['drop_prob'] = kwargs.pop('dropout')
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Dropout(dropout_layer['drop_prob']) if dropout_layer else nn.Identity()

    # 'bf16': projections as bf16x3 tcgen05 GEMMs, the core as the flash (cross) / x3 (self) tcgen05 kernels --
    # the same kernels the head engine launches; 'fp32': the exact fp32 kernels.  Shapes the tensor-core kernels
    # are not specialised for (embed_dims != 256 or 8 heads of 32) always take the fp32 kernels.
    precision = 'bf16'

    @staticmethod
    def _attn_bias(attn_mask, B, heads, Lq, Lk):
        """torch.nn.MultiheadAttention's attn_mask (bool: True = masked, or additive float; (Lq,Lk) or
        (B*heads,Lq,Lk)) -> contiguous additive fp32."""
        if attn_mask is None:
            return None
        if attn_mask.dtype == torch.bool or attn_mask.dtype == torch.uint8:
            attn_mask = torch.zeros(attn_mask.shape, dtype=torch.float32, device=attn_mask.device).masked_fill_(
                attn_mask.bool(), float('-inf'))
        attn_mask = attn_mask.float().contiguous()
        if tuple(attn_mask.shape) not in ((Lq, Lk), (B * heads, Lq, Lk)):
            raise RuntimeError(f"attn_mask shape {tuple(attn_mask.shape)} must be {(Lq, Lk)} or {(B * heads, Lq, Lk)}")
        return attn_mask

    @torch.no_grad()
    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        _require_eval(self)
        if not query.is_cuda:
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        self_attn = key is None or key is query
        same_kq_pos = key_pos is query_pos or (key_pos is None and query_pos is not None)
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None:
            if query_pos.shape == key.shape:
                key_pos = query_pos
            else:
                warnings.warn(f'position encoding of key is missing in {self.__class__.__name__}.')
        E = self.embed_dims
        if E // self.num_heads != 32:
            raise NotImplementedError("petr_b200 attention kernels are specialised for head_dim 32")
        bf = self.batch_first
        q, B, Lq = _rows(query.float(), bf)
        k, _, Lk = _rows(key.float(), bf)
        v, _, _ = _rows(value.float(), bf)
        idn, _, _ = _rows(identity.float(), bf)
        qp = _rows(query_pos.float().expand_as(query), bf)[0] if query_pos is not None else None
        kp = _rows(key_pos.float().expand_as(key), bf)[0] if key_pos is not None else None
        km = key_padding_mask.to(torch.uint8).contiguous() if key_padding_mask is not None else None
        bias = self._attn_bias(attn_mask, B, self.num_heads, Lq, Lk)
        tc_ok = self.precision == 'bf16' and E == 256 and self.num_heads == 8 and bias is None
        if tc_ok:
            out = self._forward_tc(q, k, v, idn, qp, kp, km, B, Lq, Lk, self_attn and value is key,
                                   self_attn and same_kq_pos and (kp is None) == (qp is None))
        else:
            w = self.attn.in_proj_weight.detach().float().contiguous()
            b = self.attn.in_proj_bias.detach().float().contiguous()
            Q = ops.gemm_f32(q, w[:E], b[:E], A2=qp)
            K = ops.gemm_f32(k, w[E:2 * E], b[E:2 * E], A2=kp)
            V = ops.gemm_f32(v, w[2 * E:], b[2 * E:])
            o = ops.attention_f32(Q, K, V, km, B, Lq, Lk, self.num_heads, attn_bias=bias)
            out = ops.gemm_f32(o, self.attn.out_proj.weight.detach().float().contiguous(),
                               self.attn.out_proj.bias.detach().float().contiguous(), residual=idn)
        out = out.view(B, Lq, E)
        return out if bf else out.transpose(0, 1)

    def _tc_weights(self):
        """[hi|hi|lo] / bf16 packs of the projection weights, rebuilt when a parameter's version changes."""
        sig = tuple((p.data_ptr(), p._version) for p in (self.attn.in_proj_weight, self.attn.in_proj_bias,
                                                         self.attn.out_proj.weight, self.attn.out_proj.bias))
        c = getattr(self, "_tc_cache", None)
        if c is None or c[0] != sig:
            E = self.embed_dims
            w = self.attn.in_proj_weight.detach().float()
            b = self.attn.in_proj_bias.detach().float()
            d = dict(q=ops.x3_weight(w[:E]), k=ops.x3_weight(w[E:2 * E]), v=ops.x3_weight(w[2 * E:]),
                     k16=w[E:2 * E].to(torch.bfloat16).contiguous(), v16=w[2 * E:].to(torch.bfloat16).contiguous(),
                     bq=b[:E].contiguous(), bk=b[E:2 * E].contiguous(), bv=b[2 * E:].contiguous(),
                     o=ops.x3_weight(self.attn.out_proj.weight.detach().float()),
                     bo=self.attn.out_proj.bias.detach().float().contiguous())
            self._tc_cache = c = (sig, d)
        return c[1]

    def _forward_tc(self, q, k, v, idn, qp, kp, km, B, Lq, Lk, self_attn, same_kq):
        """Tensor-core path on (B*L, E) rows.  Self-attention (short key axis, fp32-class accuracy needed on the
        residual stream): hi/lo-split x3 kernel; cross-attention: bf16 K / V and the flash kernel."""
        from .engine_tc import Q_SCALE, choose_nsplit
        T = self._tc_weights()
        g = ops.gemm_bf16_tc
        heads = self.num_heads
        dev = q.device
        o_s = torch.empty((B * Lq, 2 * self.embed_dims), dtype=torch.bfloat16, device=dev)
        qs = ops.split_rows(q, add=qp)
        if self_attn:
            ks = qs if same_kq else ops.split_rows(k, add=kp)
            Qp = g(qs, T["q"], T["bq"], x3=True, headsplit_Lk=Lq, headsplit_pair=True, name="gemm_q_side")
            Kp = g(ks, T["k"], T["bk"], x3=True, headsplit_Lk=Lk, headsplit_pair=True, name="gemm_q_side")
            Vp = g(ops.split_rows(v), T["v"], T["bv"], x3=True, headsplit_Lk=Lk, headsplit_pair=True, name="gemm_q_side")
            ops.attention_x3_tc(Qp[:, 0], Kp[:, 0], Vp[:, 0], km, B, Lq, Lk, heads, out_split=o_s, want_f32=False)
        else:
            Qh = g(qs, T["q"], T["bq"], x3=True, alpha=Q_SCALE, headsplit_Lk=Lq, name="gemm_q_side")
            kin = ops.add_rows_bf16(k, kp) if kp is not None else ops.add_rows_bf16(k, torch.zeros_like(k[:1]))
            vin = kin if (v is k and kp is None) else ops.add_rows_bf16(v, torch.zeros_like(v[:1]))
            Kh = g(kin, T["k16"], T["bk"], headsplit_Lk=Lk, name="gemm_kproj")
            Vh = g(vin, T["v16"], T["bv"], headsplit_Lk=Lk, name="gemm_vproj")
            ops.cross_attention_bf16(Qh[0], Kh[0], Vh[0], km, B, Lq, Lk, heads, choose_nsplit(B, heads, Lq, Lk),
                                     out_split=o_s, want_f32=False)
        return g(o_s, T["o"], T["bo"], residual=idn.contiguous(), x3=True, name="gemm_q_side")


@ATTENTION.register_module()
class MultiheadAttention(PETRMultiheadAttention):
    """mmcv ``MultiheadAttention`` (self-attention of the decoder layer, config :63); the
    reference's PETRMultiheadAttention is a verbatim copy of it."""


@FEEDFORWARD_NETWORK.register_module()
class FFN(BaseModule):
    """mmcv-full 1.4.0 FFN: keys ``layers.0.0.*`` (Linear E->F) and ``layers.1.*`` (Linear F->E)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2
        if act_cfg.get('type', 'ReLU') != 'ReLU':
            raise NotImplementedError("petr_b200 FFN kernels implement ReLU (all PETR configs)")
        layers = []
        c = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(c, feedforward_channels), nn.ReLU(inplace=True),
                                        nn.Dropout(ffn_drop)))
            c = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = nn.Dropout(dropout_layer['drop_prob']) if dropout_layer else nn.Identity()
        self.add_identity = add_identity
        self.embed_dims = embed_dims
        self.num_fcs = num_fcs

    precision = 'bf16'        # bf16x3 tcgen05 GEMMs (fp32-class accuracy); 'fp32': exact fp32 FFMA kernels

    @torch.no_grad()
    def forward(self, x, identity=None):
        _require_eval(self)
        if not x.is_cuda:
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        shape = x.shape
        h = x.float().contiguous().view(-1, shape[-1])
        idn = h if identity is None else identity.float().contiguous().view(-1, shape[-1])
        lins = [self.layers[i][0] for i in range(self.num_fcs - 1)] + [self.layers[self.num_fcs - 1]]
        if self.precision == 'bf16' and all(l.in_features % 64 == 0 and l.out_features % 32 == 0 for l in lins):
            sig = tuple((l.weight.data_ptr(), l.weight._version) for l in lins)
            c = getattr(self, "_tc_cache", None)
            if c is None or c[0] != sig:
                self._tc_cache = c = (sig, [ops.x3_weight(l.weight.detach().float()) for l in lins])
            hs = ops.split_rows(h)
            for i, l in enumerate(lins[:-1]):
                hs = ops.gemm_bf16_tc(hs, c[1][i], l.bias.detach().float(), act=ACT_RELU, x3=True, split_out=True,
                                      name="gemm_q_side")
            out = ops.gemm_bf16_tc(hs, c[1][-1], lins[-1].bias.detach().float(),
                                   residual=idn if self.add_identity else None, x3=True, name="gemm_q_side")
            return out.view(shape)
        for i in range(self.num_fcs - 1):
            lin = self.layers[i][0]
            h = ops.gemm_f32(h, lin.weight.detach().contiguous(), lin.bias.detach(), act=ACT_RELU)
        lin = self.layers[self.num_fcs - 1]
        out = ops.gemm_f32(h, lin.weight.detach().contiguous(), lin.bias.detach(),
                           residual=idn if self.add_identity else None)
        return out.view(shape)


def _layernorm_module(ln: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    return ops.layernorm(x.float().contiguous(), ln.weight.detach(), ln.bias.detach(), ln.eps)


@TRANSFORMER_LAYER.register_module()
class BaseTransformerLayer(BaseModule):
    """mmcv BaseTransformerLayer (the reference's vendored near-copy:
    multi_atten_decoder_layer.py:72-293)."""

    def __init__(self, attn_cfgs=None,
                 ffn_cfgs=dict(type='FFN', embed_dims=256, feedforward_channels=1024, num_fcs=2,
                               ffn_drop=0., act_cfg=dict(type='ReLU', inplace=True)),
                 operation_order=None, norm_cfg=dict(type='LN'), init_cfg=None, batch_first=False,
                 **kwargs):
        deprecated_args = dict(feedforward_channels='feedforward_channels', ffn_dropout='ffn_drop',
                               ffn_num_fcs='num_fcs')
        ffn_cfgs = copy.deepcopy(ffn_cfgs)
        for ori_name, new_name in deprecated_args.items():
            if ori_name in kwargs:
                ffn_cfgs[new_name] = kwargs[ori_name]
        super().__init__(init_cfg)
        self.batch_first = batch_first
        assert set(operation_order) & {'self_attn', 'norm', 'ffn', 'cross_attn'} == set(operation_order)
        num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        assert num_attn == len(attn_cfgs)
        self.num_attn = num_attn
        self.operation_order = operation_order
        self.norm_cfg = norm_cfg
        self.pre_norm = operation_order[0] == 'norm'
        self.attentions = nn.ModuleList()
        index = 0
        for op in operation_order:
            if op in ('self_attn', 'cross_attn'):
                cfg = copy.deepcopy(attn_cfgs[index])
                if 'batch_first' in cfg:
                    assert self.batch_first == cfg['batch_first']
                else:
                    cfg['batch_first'] = self.batch_first
                attention = build_attention(cfg)
                attention.operation_name = op
                self.attentions.append(attention)
                index += 1
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = nn.ModuleList()
        num_ffns = operation_order.count('ffn')
        if isinstance(ffn_cfgs, dict):
            ffn_cfgs = [copy.deepcopy(ffn_cfgs) for _ in range(num_ffns)]
        for i in range(num_ffns):
            if 'embed_dims' not in ffn_cfgs[i]:
                ffn_cfgs[i]['embed_dims'] = self.embed_dims
            else:
                assert ffn_cfgs[i]['embed_dims'] == self.embed_dims
            self.ffns.append(build_feedforward_network(ffn_cfgs[i]))
        self.norms = nn.ModuleList()
        if norm_cfg.get('type', 'LN') != 'LN':
            raise NotImplementedError("petr_b200 implements LayerNorm (norm_cfg type 'LN')")
        for _ in range(operation_order.count('norm')):
            self.norms.append(nn.LayerNorm(self.embed_dims, eps=norm_cfg.get('eps', 1e-5)))

    @torch.no_grad()
    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        _require_eval(self)
        norm_index = attn_index = ffn_index = 0
        identity = query
        # multi_atten_decoder_layer.py:215-232
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
            warnings.warn(f'Use same attn_mask in all attentions in {self.__class__.__name__} ')
        else:
            assert len(attn_masks) == self.num_attn, f'The length of attn_masks {len(attn_masks)} must be equal ' \
                f'to the number of attention in operation_order {self.num_attn}'
        for layer in self.operation_order:
            if layer == 'self_attn':
                query = self.attentions[attn_index](
                    query, query, query, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=query_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=query_key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'norm':
                query = _layernorm_module(self.norms[norm_index], query)
                norm_index += 1
            elif layer == 'cross_attn':
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=key_pos, attn_mask=attn_masks[attn_index], key_padding_mask=key_padding_mask,
                    **kwargs)
                attn_index += 1
                identity = query
            elif layer == 'ffn':
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query


@TRANSFORMER_LAYER.register_module()
class PETRTransformerDecoderLayer(BaseTransformerLayer):
    """petr_transformer.py:113-224.  ``with_cp`` (activation checkpointing) only matters in
    training and is accepted for config compatibility."""

    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), ffn_num_fcs=2,
                 with_cp=True, **kwargs):
        super().__init__(attn_cfgs=attn_cfgs, feedforward_channels=feedforward_channels,
                         ffn_dropout=ffn_dropout, operation_order=operation_order, act_cfg=act_cfg,
                         norm_cfg=norm_cfg, ffn_num_fcs=ffn_num_fcs, **kwargs)
        assert len(operation_order) == 6
        assert set(operation_order) == set(['self_attn', 'norm', 'cross_attn', 'ffn'])
        self.use_checkpoint = with_cp


class TransformerLayerSequence(BaseModule):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        else:
            assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = nn.ModuleList([build_transformer_layer(c) for c in transformerlayers])
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm

    def forward(self, query, key, value, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        for layer in self.layers:
            query = layer(query, key, value, query_pos=query_pos, key_pos=key_pos,
                          attn_masks=attn_masks, query_key_padding_mask=query_key_padding_mask,
                          key_padding_mask=key_padding_mask, **kwargs)
        return query


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class PETRTransformerEncoder(TransformerLayerSequence):
    """petr_transformer.py:371-397 (unused by every shipped config)."""

    def __init__(self, *args, post_norm_cfg=dict(type='LN'), **kwargs):
        super().__init__(*args, **kwargs)
        if post_norm_cfg is not None:
            self.post_norm = nn.LayerNorm(self.embed_dims) if self.pre_norm else None
        else:
            assert not self.pre_norm
            self.post_norm = None

    def forward(self, *args, **kwargs):
        x = super().forward(*args, **kwargs)
        if self.post_norm is not None:
            x = _layernorm_module(self.post_norm, x)
        return x


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class PETRTransformerDecoder(TransformerLayerSequence):
    """petr_transformer.py:401-447."""

    def __init__(self, *args, post_norm_cfg=dict(type='LN'), return_intermediate=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.return_intermediate = return_intermediate
        self.post_norm = nn.LayerNorm(self.embed_dims) if post_norm_cfg is not None else None

    @torch.no_grad()
    def forward(self, query, *args, **kwargs):
        kwargs.pop('reg_branch', None)
        if not self.return_intermediate:
            x = super().forward(query, *args, **kwargs)
            if self.post_norm:
                x = _layernorm_module(self.post_norm, x)[None]
            return x
        intermediate = []
        for layer in self.layers:
            query = layer(query, *args, **kwargs)
            intermediate.append(_layernorm_module(self.post_norm, query)
                                if self.post_norm is not None else query)
        return torch.stack(intermediate)


@TRANSFORMER.register_module()
class PETRTransformer(BaseModule):
    """petr_transformer.py:34-109."""

    def __init__(self, encoder=None, decoder=None, init_cfg=None, cross=False):
        super().__init__(init_cfg=init_cfg)
        self.encoder = build_transformer_layer_sequence(encoder) if encoder is not None else None
        self.decoder = build_transformer_layer_sequence(decoder)
        self.embed_dims = self.decoder.embed_dims
        self.cross = cross

    def init_weights(self):
        # xavier-uniform on every >1-D weight, zero bias (petr_transformer.py:62-67)
        for m in self.modules():
            if hasattr(m, 'weight') and m.weight is not None and m.weight.dim() > 1:
                nn.init.xavier_uniform_(m.weight)
                if getattr(m, 'bias', None) is not None:
                    nn.init.constant_(m.bias, 0)
        self._is_init = True

    @torch.no_grad()
    def forward(self, x, mask, query_embed, pos_embed, reg_branch=None):
        bs, n, c, h, w = x.shape
        memory = x.permute(1, 3, 4, 0, 2).reshape(-1, bs, c)
        pos_embed = pos_embed.permute(1, 3, 4, 0, 2).reshape(-1, bs, c)
        query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)
        mask = mask.view(bs, -1)
        target = torch.zeros_like(query_embed)
        out_dec = self.decoder(query=target, key=memory, value=memory, key_pos=pos_embed,
                               query_pos=query_embed, key_padding_mask=mask, reg_branch=reg_branch)
        out_dec = out_dec.transpose(1, 2)
        memory = memory.reshape(n, h, w, bs, c).permute(3, 0, 4, 1, 2)
        return out_dec, memory


# ------------------------------------------------------------------------------------------
# positional encodings
# ------------------------------------------------------------------------------------------
@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding3D(BaseModule):
    """positional_encoding.py:15-110."""
    kind = "sine3d"

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi, eps=1e-6,
                 offset=0., init_cfg=None):
        super().__init__(init_cfg)
        if normalize:
            assert isinstance(scale, (float, int)), 'when normalize is set, scale should be ' \
                f'provided and in float or int type, found {type(scale)}'
        self.num_feats = num_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = scale
        self.eps = eps
        self.offset = offset

    @torch.no_grad()
    def forward(self, mask):
        if not mask.is_cuda:
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        from .engine import sine_dim_t
        B, N, H, W = mask.shape
        dim_t = sine_dim_t(self.num_feats, self.temperature).to(mask.device)
        tok = ops.sine3d_features(mask.to(torch.uint8).contiguous(), dim_t, self.num_feats,
                                  self.normalize, self.scale, self.eps, self.offset)
        return tok.view(B, N, H, W, 3 * self.num_feats).permute(0, 1, 4, 2, 3)

    def __repr__(self):
        return (f'{self.__class__.__name__}(num_feats={self.num_feats}, temperature={self.temperature}, '
                f'normalize={self.normalize}, scale={self.scale}, eps={self.eps})')


@POSITIONAL_ENCODING.register_module()
class LearnedPositionalEncoding3D(BaseModule):
    """positional_encoding.py:114-167 (unused by every shipped config): parameter container
    kept for config / checkpoint compatibility."""

    def __init__(self, num_feats, row_num_embed=50, col_num_embed=50,
                 init_cfg=dict(type='Uniform', layer='Embedding')):
        super().__init__(init_cfg)
        self.row_embed = nn.Embedding(row_num_embed, num_feats)
        self.col_embed = nn.Embedding(col_num_embed, num_feats)
        self.num_feats = num_feats
        self.row_num_embed = row_num_embed
        self.col_num_embed = col_num_embed

    kind = "learned"

    @torch.no_grad()
    def forward(self, mask):
        """positional_encoding.py:137-160: mask (bs, h, w) -> (bs, 2*num_feats, h, w) = [col_embed(x) | row_embed(y)]."""
        if not mask.is_cuda:
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        G = int(mask.shape[0])
        h, w = mask.shape[-2:]
        tok = ops.learned_pos_features(self.row_embed.weight.detach().float().contiguous(),
                                       self.col_embed.weight.detach().float().contiguous(), G, h, w)
        return ops.tokens_to_nchw(tok, G, h * w).view(G, 2 * self.num_feats, h, w)

    def __repr__(self):
        return (f'{self.__class__.__name__}(num_feats={self.num_feats}, row_num_embed={self.row_num_embed}, '
                f'col_num_embed={self.col_num_embed})')


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding(BaseModule):
    """mmdet 2.24.1 `SinePositionalEncoding` (un-vendored): the 2-D encoding the head's default
    `positional_encoding` dict names (petr_head.py:96-99) and the with_multiview=False branch evaluates per camera
    (petr_head.py:403-409).  mask (bs, h, w) -> (bs, 2*num_feats, h, w) = [pos_y | pos_x], sin/cos interleaved."""
    kind = "sine2d"

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi, eps=1e-6,
                 offset=0., init_cfg=None):
        super().__init__(init_cfg)
        if normalize:
            assert isinstance(scale, (float, int)), 'when normalize is set, scale should be ' \
                f'provided and in float or int type, found {type(scale)}'
        self.num_feats = num_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = scale
        self.eps = eps
        self.offset = offset

    @torch.no_grad()
    def forward(self, mask):
        if not mask.is_cuda:
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        from .engine import sine_dim_t
        G, H, W = mask.shape
        dim_t = sine_dim_t(self.num_feats, self.temperature).to(mask.device)
        tok = ops.sine2d_features(mask.to(torch.uint8).contiguous(), dim_t, self.num_feats, self.normalize,
                                  self.scale, self.eps, self.offset)
        return ops.tokens_to_nchw(tok, G, H * W).view(G, 2 * self.num_feats, H, W)

    def __repr__(self):
        return (f'{self.__class__.__name__}(num_feats={self.num_feats}, temperature={self.temperature}, '
                f'normalize={self.normalize}, scale={self.scale}, eps={self.eps})')


# ------------------------------------------------------------------------------------------
# bbox coder (pc_range holder; decode is the first 'next' row of SURVEY.md section 8f)
# ------------------------------------------------------------------------------------------
@BBOX_CODERS.register_module()
class NMSFreeCoder:
    """nms_free_coder.py:18-119.  Holds pc_range for the head and decodes the last decoder layer's
    predictions: sigmoid -> top-`max_num` over queries x classes -> denormalise -> post_center_range
    (and score) filter -- ONE kernel launch for the whole batch (petr_nms_free_decode).
    CUDA tensors only: like every module of this package there is no CPU / PyTorch fallback."""

    def __init__(self, pc_range, voxel_size=None, post_center_range=None, max_num=100,
                 score_threshold=None, num_classes=10):
        self.pc_range = pc_range
        self.voxel_size = voxel_size
        self.post_center_range = post_center_range
        self.max_num = max_num
        self.score_threshold = score_threshold
        self.num_classes = num_classes

    def encode(self):
        pass

    def decode_padded(self, cls_scores: torch.Tensor, bbox_preds: torch.Tensor, bottom_center: bool = False):
        """Graph-capturable form: fixed-shape outputs (boxes (B,max_num,9), scores, labels, query index, keep-mask,
        kept-count), rows sorted by descending score; no host synchronisation."""
        if self.post_center_range is None:
            raise NotImplementedError('Need to reorganize output as a batch, only support '
                                      'post_center_range is not None for now!')
        if not cls_scores.is_cuda:
            raise RuntimeError("petr_b200: NMSFreeCoder.decode needs CUDA tensors (no CPU fallback)")
        if cls_scores.shape[-1] != self.num_classes:
            raise ValueError(f"cls_scores has {cls_scores.shape[-1]} classes, coder was built for {self.num_classes}")
        return ops.nms_free_decode(cls_scores, bbox_preds, self.max_num, self.post_center_range,
                                   self.score_threshold if self.score_threshold else None, bottom_center)

    def _decode_batch(self, cls_scores: torch.Tensor, bbox_preds: torch.Tensor, bottom_center: bool = False):
        """cls_scores (B, Q, C), bbox_preds (B, Q, code) -> list of dicts (variable length per sample, as the
        reference returns them: the boolean selection below is the one data-dependent step)."""
        boxes, scores, labels, _, mask, _ = self.decode_padded(cls_scores, bbox_preds, bottom_center)
        return [dict(bboxes=boxes[b][mask[b]], scores=scores[b][mask[b]], labels=labels[b][mask[b]])
                for b in range(boxes.shape[0])]

    def decode_single(self, cls_scores, bbox_preds):
        return self._decode_batch(cls_scores[None], bbox_preds[None])[0]

    def decode(self, preds_dicts):
        return self._decode_batch(preds_dicts['all_cls_scores'][-1], preds_dicts['all_bbox_preds'][-1])


# ------------------------------------------------------------------------------------------
# heads
# ------------------------------------------------------------------------------------------
class NormedLinear(nn.Linear):
    """mmdet 2.24.1 `NormedLinear` (un-vendored; petr_head.py:231-232 when normedlinear=True): weight rows and the
    activations are divided by (their L2 norm over dim 1)^power + eps, the activations multiplied by the temperature
    (the published attribute name `tempearture` is kept).  Parameter container + standalone forward; inside the head
    the engine launches the same kernels."""

    def __init__(self, *args, tempearture=20, power=1.0, eps=1e-6, **kwargs):
        super().__init__(*args, **kwargs)
        self.tempearture = tempearture
        self.power = power
        self.eps = eps
        nn.init.normal_(self.weight, mean=0, std=0.01)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        w = self.weight.detach().float().contiguous()
        w = ops.norm_scale(w, 1, w.shape[0], False, self.power, self.eps, 1.0)
        xf = x.float().contiguous()
        if xf.dim() == 2:       # norm(dim=1) of a 2-D input is per row
            xs = ops.norm_scale(xf, 1, xf.shape[0], False, self.power, self.eps, self.tempearture)
        else:                   # ... of an N-D input it runs over axis 1
            Bn, Q = xf.shape[0], xf.shape[1]
            inner = xf.numel() // (Bn * Q)
            xs = ops.norm_scale(xf.view(Bn, Q, inner), Bn, Q, True, self.power, self.eps, self.tempearture)
        out = ops.gemm_f32(xs.view(-1, xf.shape[-1]), w, self.bias.detach().float() if self.bias is not None else None)
        return out.view(*x.shape[:-1], self.out_features)


class _LossCfg:
    """Holds a loss config; only ``use_sigmoid`` is needed by the forward (petr_head.py:200)."""

    def __init__(self, cfg):
        self.cfg = copy.deepcopy(cfg)
        self.use_sigmoid = bool(cfg.get('use_sigmoid', False)) if cfg else False


class _PETRHeadBase(BaseModule):
    _version = 2
    _optional_transformer = False

    def _common_init(self, num_classes, in_channels, num_query, num_reg_fcs, transformer,
                     sync_cls_avg_factor, positional_encoding, code_weights, bbox_coder, loss_cls,
                     loss_bbox, loss_iou, train_cfg, test_cfg, with_position, with_multiview,
                     depth_step, depth_num, LID, depth_start, position_range, normedlinear, kwargs):
        self.code_size = kwargs.get('code_size', 10)
        cw = code_weights if code_weights is not None else [1.0] * 8 + [0.2, 0.2]
        self._code_weights_list = list(cw)[:self.code_size]
        self.bg_cls_weight = 0
        self.sync_cls_avg_factor = sync_cls_avg_factor
        self.num_query = num_query
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.num_reg_fcs = num_reg_fcs
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.fp16_enabled = False
        self.embed_dims = 256
        self.depth_step = depth_step
        self.depth_num = depth_num
        self.position_dim = 3 * self.depth_num
        self.position_range = position_range
        self.LID = LID
        self.depth_start = depth_start
        self.with_position = with_position
        self.with_multiview = with_multiview
        assert 'num_feats' in positional_encoding
        num_feats = positional_encoding['num_feats']
        assert num_feats * 2 == self.embed_dims, 'embed_dims should be exactly 2 times of ' \
            f'num_feats. Found {self.embed_dims} and {num_feats}.'
        self.act_cfg = transformer.get('act_cfg', dict(type='ReLU', inplace=True))
        self.num_pred = 6
        self.normedlinear = normedlinear
        self.loss_cls = _LossCfg(loss_cls)
        self.loss_bbox = _LossCfg(loss_bbox)
        self.loss_iou = _LossCfg(loss_iou)
        self.cls_out_channels = num_classes if self.loss_cls.use_sigmoid else num_classes + 1
        self.positional_encoding = build_positional_encoding(positional_encoding)
        if transformer is not None and TRANSFORMER.get(transformer.get('type')) is None and self._optional_transformer:
            self.transformer, self.transformer_cfg = None, copy.deepcopy(transformer)    # DepthrHead, see there
        else:
            self.transformer = build_transformer(transformer)
        self.code_weights = nn.Parameter(torch.tensor(self._code_weights_list, requires_grad=False),
                                         requires_grad=False)
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.pc_range = self.bbox_coder.pc_range
        self._engine = None
        self._train_engine = None
        self._loss_impl = None
        self.precision = 'bf16'
        # None: samples are simply data-parallel.  dict(world, rank, group): this process computes its
        # share of a key-split over camera views (every rank passes the full inputs, gets the full outputs)
        self.keysplit = None
        # True: the whole forward is captured once per (input buffer, shape) as a CUDA graph and
        # replayed; the returned tensors are then the graph's static outputs (valid until the next call)
        self.use_cuda_graph = False

    # -- layers ---------------------------------------------------------------------------
    def _make_cls_branch(self):
        layers = []
        for _ in range(self.num_reg_fcs):
            layers += [nn.Linear(self.embed_dims, self.embed_dims), nn.LayerNorm(self.embed_dims),
                       nn.ReLU(inplace=True)]
        if self.normedlinear:
            layers.append(NormedLinear(self.embed_dims, self.cls_out_channels))
        else:
            layers.append(nn.Linear(self.embed_dims, self.cls_out_channels))
        return nn.Sequential(*layers)

    def _make_reg_branch(self):
        layers = []
        for _ in range(self.num_reg_fcs):
            layers += [nn.Linear(self.embed_dims, self.embed_dims), nn.ReLU()]
        layers.append(nn.Linear(self.embed_dims, self.code_size))
        return nn.Sequential(*layers)

    def _init_common_layers(self):
        E = self.embed_dims
        self.input_proj = nn.Conv2d(self.in_channels, E, kernel_size=1)
        if self.with_multiview:
            self.adapt_pos3d = nn.Sequential(nn.Conv2d(E * 3 // 2, E * 4, 1), nn.ReLU(),
                                             nn.Conv2d(E * 4, E, 1))
        else:
            self.adapt_pos3d = nn.Sequential(nn.Conv2d(E, E, 1), nn.ReLU(), nn.Conv2d(E, E, 1))
        if self.with_position:
            self.position_encoder = nn.Sequential(nn.Conv2d(self.position_dim, E * 4, 1), nn.ReLU(),
                                                  nn.Conv2d(E * 4, E, 1))
        self.reference_points = nn.Embedding(self.num_query, 3)
        self.query_embedding = nn.Sequential(nn.Linear(E * 3 // 2, E), nn.ReLU(), nn.Linear(E, E))

    def init_weights(self):
        """petr_head.py:276-284."""
        if self.transformer is not None:
            self.transformer.init_weights()
        nn.init.uniform_(self.reference_points.weight.data, 0, 1)
        if self.loss_cls.use_sigmoid:
            bias_init = float(-math.log((1 - 0.01) / 0.01))
            for m in self.cls_branches:
                nn.init.constant_(m[-1].bias, bias_init)
        self._invalidate_engine()
        self._is_init = True

    # -- engine cache ---------------------------------------------------------------------
    @property
    def engine(self) -> HeadEngine:
        """The engine keeps derived copies of the parameters (bf16 / [hi|hi|lo] packs, concatenated
        K/V weights) and captured CUDA graphs.  They are dropped whenever the parameters may have
        changed: load_state_dict / _load_from_state_dict (also reached by a detector-level or mmcv
        `load_checkpoint` recursion), _apply (.to/.cuda/.half), init_weights, and -- checked on every
        access -- any in-place edit, through the tensors' version counters."""
        if self._engine is None:
            self._engine = HeadEngine(self)
            self._engine_sig = self._param_signature()
        else:
            sig = self._param_signature()
            if sig != self._engine_sig:
                self._engine.invalidate()
                self._engine_sig = sig
        return self._engine

    def _param_signature(self):
        v = n = 0
        for p in self.parameters():
            v += p._version
            n += 1
        return (v, n)

    def _invalidate_engine(self):
        if self._engine is not None:
            self._engine.invalidate()
            self._engine_sig = self._param_signature()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate_engine()
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._invalidate_engine()
        return out

    # -- hot path --------------------------------------------------------------------------
    @torch.no_grad()
    def position_embeding(self, img_feats, img_metas, masks=None):
        """petr_head.py:286-334.  fp32 path.  Returns ((B,N,256,H,W) fp32, bool (B,N,H,W))."""
        _require_eval(self)
        feat = img_feats[self.position_level]
        if not feat.is_cuda:
            raise RuntimeError("petr_b200: CUDA tensors required (no CPU fallback)")
        B, N, C, H, W = feat.shape
        eng = self.engine
        m8 = masks.to(torch.uint8).contiguous() if masks is not None else None
        mi = eng.meta_inputs(img_metas, N, H, W, feat.device)
        pos, cmask = eng.position_embedding_tokens(eng.packed(feat.device), mi, B, N, H, W, m8)
        pos = pos.view(B, N, H, W, self.embed_dims).permute(0, 1, 4, 2, 3)
        return pos, cmask.to(torch.bool)

    def forward(self, mlvl_feats, img_metas):
        """petr_head.py:366-468 / petrv2_head.py:429-540.

        eval mode: the inference engine (no autograd).  With `use_cuda_graph = True` the returned tensors are the
        static outputs of the captured graph: valid until the next call on this head (clone to keep), and the feature
        tensor must be the same device buffer every call (see HeadEngine._graph_forward).  Not thread-safe: one head
        object per host thread.
        train mode (`.train()`): the differentiable forward of petr_b200/train.py (fp32 kernels, dropout active);
        follow with `loss(...)` and `.backward()` as petr3d.py:126-128 does."""
        x = mlvl_feats[self.position_level]
        if self.training:
            return self.forward_autograd(mlvl_feats, img_metas, dropout=True)
        with torch.no_grad():
            return self.engine.forward(x, img_metas, precision=self.precision, keysplit=self.keysplit,
                                       use_graph=self.use_cuda_graph)

    def forward_autograd(self, mlvl_feats, img_metas, dropout: bool = True):
        """Differentiable forward (what `.train()` + forward runs); `dropout=False` gives the eval-mode function with
        gradients (parity tests against the reference's autograd)."""
        from .train import TrainEngine
        if self._train_engine is None:
            self._train_engine = TrainEngine(self)
        return self._train_engine.forward(mlvl_feats[self.position_level], img_metas, training=dropout)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        """Legacy-key remap of petr_head.py:336-364 (version < 2 checkpoints)."""
        version = local_metadata.get('version', None)
        if (version is None or version < 2) and self._remap_legacy:
            convert = {'.self_attn.': '.attentions.0.', '.multihead_attn.': '.attentions.1.',
                       '.decoder.norm.': '.decoder.post_norm.'}
            for k in list(state_dict.keys()):
                for ori, new in convert.items():
                    if ori in k:
                        state_dict[k.replace(ori, new)] = state_dict.pop(k)
                        break
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)
        self._invalidate_engine()

    def get_bboxes(self, preds_dicts, img_metas, rescale=False):
        """petr_head.py:731-751: NMS-free decode of the last layer, gravity centre -> bottom centre,
        wrapped in each sample's `box_type_3d` (left as a plain tensor when img_metas carries none,
        e.g. without mmdet3d)."""
        preds = self.bbox_coder._decode_batch(preds_dicts['all_cls_scores'][-1], preds_dicts['all_bbox_preds'][-1],
                                              bottom_center=True)          # z -= h / 2 inside the kernel
        ret_list = []
        for i, pr in enumerate(preds):
            bboxes = pr['bboxes']
            box_type = img_metas[i].get('box_type_3d') if img_metas is not None else None
            if box_type is not None:
                bboxes = box_type(bboxes, bboxes.size(-1))
            ret_list.append([bboxes, pr['scores'], pr['labels']])
        return ret_list

    def loss(self, gt_bboxes_list, gt_labels_list, preds_dicts, gt_bboxes_ignore=None):
        """petr_head.py:647-728: dict(loss_cls, loss_bbox, d0.loss_cls, ... d4.loss_bbox).  Focal / L1 loss kernels on
        Hungarian-matched queries; the cost matrix is a kernel, the assignment scipy on the host as in the reference."""
        from .train import HeadLoss
        if self._loss_impl is None:
            self._loss_impl = HeadLoss(self)
        return self._loss_impl(gt_bboxes_list, gt_labels_list, preds_dicts, gt_bboxes_ignore)


_HEAD_DEFAULTS = dict(
    num_query=100, num_reg_fcs=2, transformer=None, sync_cls_avg_factor=False,
    positional_encoding=dict(type='SinePositionalEncoding', num_feats=128, normalize=True),
    code_weights=None, bbox_coder=None,
    loss_cls=dict(type='CrossEntropyLoss', bg_cls_weight=0.1, use_sigmoid=False, loss_weight=1.0,
                  class_weight=1.0),
    loss_bbox=dict(type='L1Loss', loss_weight=5.0), loss_iou=dict(type='GIoULoss', loss_weight=2.0),
    train_cfg=None, test_cfg=dict(max_per_img=100), with_position=True, with_multiview=False,
    depth_step=0.8, depth_num=64, LID=False, depth_start=1,
    position_range=[-65, -65, -8.0, 65, 65, 8.0], init_cfg=None, normedlinear=False)


@HEADS.register_module()
class PETRHead(_PETRHeadBase):
    """petr_head.py:47-468.  The six cls/reg branches are one shared module (state_dict holds
    six aliased copies), ``position_level`` is 0."""
    _remap_legacy = True

    def __init__(self, num_classes, in_channels, **kw):
        a = dict(_HEAD_DEFAULTS)
        extra = {k: kw.pop(k) for k in list(kw) if k not in a}
        a.update(kw)
        super().__init__(a.pop('init_cfg'))
        a.pop('train_cfg_unused', None)
        self.position_level = 0
        self._common_init(num_classes, in_channels, a['num_query'], a['num_reg_fcs'], a['transformer'],
                          a['sync_cls_avg_factor'], a['positional_encoding'], a['code_weights'],
                          a['bbox_coder'], a['loss_cls'], a['loss_bbox'], a['loss_iou'], a['train_cfg'],
                          a['test_cfg'], a['with_position'], a['with_multiview'], a['depth_step'],
                          a['depth_num'], a['LID'], a['depth_start'], a['position_range'],
                          a['normedlinear'], extra)
        self._init_layers()

    def _init_layers(self):
        self._init_common_layers()
        fc_cls = self._make_cls_branch()
        reg_branch = self._make_reg_branch()
        self.cls_branches = nn.ModuleList([fc_cls for _ in range(self.num_pred)])
        self.reg_branches = nn.ModuleList([reg_branch for _ in range(self.num_pred)])


class SELayer(nn.Module):
    """petrv2_head.py:48-60 (parameter container; the gate runs as a GEMM epilogue)."""

    def __init__(self, channels):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, channels, 1, bias=True)
        self.act1 = nn.ReLU()
        self.conv_expand = nn.Conv2d(channels, channels, 1, bias=True)
        self.gate = nn.Sigmoid()


class RegLayer(nn.Module):
    """petrv2_head.py:63-95 (parameter container)."""

    def __init__(self, embed_dims=256, shared_reg_fcs=2, group_reg_dims=(2, 1, 3, 2, 2), drop=0.0):
        super().__init__()
        layers = []
        for _ in range(shared_reg_fcs):
            layers += [nn.Linear(embed_dims, embed_dims), nn.ReLU(), nn.Dropout(drop)]
        self.reg_branch = nn.Sequential(*layers)
        self.task_heads = nn.ModuleList([
            nn.Sequential(nn.Linear(embed_dims, embed_dims), nn.ReLU(), nn.Linear(embed_dims, d))
            for d in group_reg_dims])


@HEADS.register_module()
class PETRv2Head(_PETRHeadBase):
    """petrv2_head.py:99-540: feature-guided position encoder (SELayer), multi-task RegLayer,
    time-normalised velocity, configurable position_level, unshared (deep-copied) branches."""
    _remap_legacy = True      # petrv2_head.py:399-421

    def __init__(self, num_classes, in_channels, position_level=0, group_reg_dims=(2, 1, 3, 2, 2),
                 with_fpe=False, with_time=False, with_multi=False, **kw):
        a = dict(_HEAD_DEFAULTS)
        extra = {k: kw.pop(k) for k in list(kw) if k not in a}
        a.update(kw)
        super().__init__(a.pop('init_cfg'))
        self.position_level = position_level
        self.group_reg_dims = group_reg_dims
        self.with_fpe = with_fpe
        self.with_time = with_time
        self.with_multi = with_multi
        self._common_init(num_classes, in_channels, a['num_query'], a['num_reg_fcs'], a['transformer'],
                          a['sync_cls_avg_factor'], a['positional_encoding'], a['code_weights'],
                          a['bbox_coder'], a['loss_cls'], a['loss_bbox'], a['loss_iou'], a['train_cfg'],
                          a['test_cfg'], a['with_position'], a['with_multiview'], a['depth_step'],
                          a['depth_num'], a['LID'], a['depth_start'], a['position_range'],
                          a['normedlinear'], extra)
        self._init_layers()

    def _init_layers(self):
        self._init_common_layers()
        fc_cls = self._make_cls_branch()
        if self.with_multi:
            reg_branch = RegLayer(self.embed_dims, self.num_reg_fcs, self.group_reg_dims)
        else:
            reg_branch = self._make_reg_branch()
        self.cls_branches = nn.ModuleList([copy.deepcopy(fc_cls) for _ in range(self.num_pred)])
        self.reg_branches = nn.ModuleList([copy.deepcopy(reg_branch) for _ in range(self.num_pred)])
        if self.with_fpe:
            self.fpe = SELayer(self.embed_dims)


@HEADS.register_module()
class DepthrHead(PETRHead):
    """depthr_head.py:65-1001, the part SURVEY.md section 8f row 4 names: `position_embeding` (:338-386, identical
    arithmetic to PETRHead's) on the same kernels, with the reference's constructor kwargs and the state_dict keys
    of the layers it shares with PETRHead (input_proj, adapt_pos3d, position_encoder, reference_points,
    query_embedding, cls / reg branches).  The depth-specific transformer (`DepthrTransformer`,
    `MultiAttentionDecoderLayer`) and `DepthGTEncoder` belong to a model family outside this path: their configs
    are kept (`transformer_cfg`, `depth_gt_encoder_cfg`) and `forward` says so instead of running."""
    _remap_legacy = True
    _optional_transformer = True

    def __init__(self, num_classes, in_channels, depth_gt_encoder=None, **kw):
        self.depth_gt_encoder_cfg = copy.deepcopy(depth_gt_encoder)
        self.depth_gt_encoder = None
        super().__init__(num_classes, in_channels, **kw)

    def forward(self, mlvl_feats, img_metas, gt_bboxes_3d=None):
        if self.transformer is None or self.depth_gt_encoder_cfg is not None:
            raise NotImplementedError(
                "DepthrHead.forward needs DepthrTransformer / DepthGTEncoder, a model family outside the PETR head hot "
                "path (DESIGN.md section 8); position_embeding() is the shared piece and runs on the CUDA path")
        return super().forward(mlvl_feats, img_metas)
