"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference head from /root/reference.

The reference (sty61010/PETR) is a pure-Python mmdetection3d plugin whose third-party
dependencies (mmcv-full 1.4.0, mmdet 2.24.1, mmdet3d 0.17.1; requirements.txt:13-14,
install.md:24-32) are not installed in this image.  This module registers small stand-in
modules for those packages in ``sys.modules`` (behaviour restated from the published
mmcv/mmdet semantics; where the reference vendors a copy the stub follows that copy:
petr_transformer.py:228-367 == mmcv MultiheadAttention, multi_atten_decoder_layer.py:72-293
== mmcv BaseTransformerLayer, utils/detr.py:15-31 == mmdet inverse_sigmoid) and then loads
the reference's own hot-path files BY PATH so they execute unmodified.

It is used in THIS container only, to (a) pin ``oracle/petr_oracle.py`` (the travelling
restatement) against the real reference and (b) generate ``tests/golden/*.npz`` through
``oracle/make_golden.py``.  /root/reference does not exist on the GPU box; nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file.
"""
from __future__ import annotations

import copy
import importlib.util
import math
import os
import sys
import types
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("PETR_REFERENCE_ROOT", "/root/reference")
_PLUGIN = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin")


def available() -> bool:
    return os.path.isfile(os.path.join(_PLUGIN, "models", "dense_heads", "petr_head.py"))


# --------------------------------------------------------------------------------------
# stand-ins for mmcv / mmdet / mmdet3d
# --------------------------------------------------------------------------------------
class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        cfg = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        typ = cfg.pop("type")
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()
        self._is_init = True


class _ModuleList(_BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        _BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class _Sequential(_BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        _BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def _identity_decorator(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def deco(fn):
        return fn
    return deco


def _xavier_init(module, gain=1, bias=0, distribution="normal"):
    if hasattr(module, "weight") and module.weight is not None:
        if distribution == "uniform":
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _constant_init(module, val, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _bias_init_with_prob(p):
    return float(-math.log((1 - p) / p))


def _inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def _build_norm_layer(cfg, num_features, postfix=""):
    assert cfg["type"] == "LN"
    return "ln" + str(postfix), nn.LayerNorm(num_features, eps=cfg.get("eps", 1e-5))


def _build_dropout(cfg, default_args=None):
    assert cfg["type"] == "Dropout"
    return nn.Dropout(cfg.get("drop_prob", 0.5))


def _build_activation_layer(cfg):
    assert cfg["type"] == "ReLU"
    return nn.ReLU(inplace=cfg.get("inplace", False))


def _install_stubs():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_petr_oracle_stub", False):
        return sys.modules["mmcv"]._registries

    def mod(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], child, m)
        return m

    R = dict(
        ATTENTION=_Registry("attention"),
        TRANSFORMER_LAYER=_Registry("transformerLayer"),
        TRANSFORMER_LAYER_SEQUENCE=_Registry("transformer-layers sequence"),
        POSITIONAL_ENCODING=_Registry("position encoding"),
        FEEDFORWARD_NETWORK=_Registry("feed-forward Network"),
        TRANSFORMER=_Registry("Transformer"),
        HEADS=_Registry("head"),
        LOSSES=_Registry("loss"),
        BBOX_CODERS=_Registry("bbox_coder"),
        NECKS=_Registry("neck"),
    )

    mmcv = mod("mmcv")
    mmcv._petr_oracle_stub = True
    mmcv._registries = R
    cnn = mod("mmcv.cnn")
    bricks = mod("mmcv.cnn.bricks")
    registry = mod("mmcv.cnn.bricks.registry")
    tfm = mod("mmcv.cnn.bricks.transformer")
    drop = mod("mmcv.cnn.bricks.drop")
    runner = mod("mmcv.runner")
    base_module = mod("mmcv.runner.base_module")
    utils = mod("mmcv.utils")

    class ConvModule(nn.Module):
        """mmcv ConvModule restricted to what cp_fpn.py builds in every PETR config (norm_cfg = act_cfg = None):
        a bare nn.Conv2d held as `.conv`."""

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None,
                     act_cfg=None, inplace=True, **kwargs):
            super().__init__()
            assert conv_cfg is None and norm_cfg is None and act_cfg is None
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding)

        def forward(self, x):
            return self.conv(x)

    cnn.ConvModule = ConvModule
    cnn.Conv2d = nn.Conv2d
    cnn.Linear = nn.Linear
    cnn.build_activation_layer = _build_activation_layer
    cnn.bias_init_with_prob = _bias_init_with_prob
    cnn.xavier_init = _xavier_init
    cnn.constant_init = _constant_init
    cnn.kaiming_init = _kaiming_init
    cnn.build_norm_layer = _build_norm_layer
    for k in ("ATTENTION", "TRANSFORMER_LAYER", "TRANSFORMER_LAYER_SEQUENCE",
              "POSITIONAL_ENCODING", "FEEDFORWARD_NETWORK"):
        setattr(registry, k, R[k])
    drop.build_dropout = _build_dropout
    runner.BaseModule = _BaseModule
    runner.ModuleList = _ModuleList
    runner.Sequential = _Sequential
    runner.force_fp32 = _identity_decorator
    runner.auto_fp16 = _identity_decorator
    base_module.BaseModule = _BaseModule
    base_module.ModuleList = _ModuleList
    base_module.Sequential = _Sequential
    utils.deprecated_api_warning = _identity_decorator

    # ---- mmcv.cnn.bricks.transformer ------------------------------------------------
    class FFN(_BaseModule):
        """mmcv 1.4.0 FFN: Sequential(Sequential(Linear, act, Dropout) x (num_fcs-1), Linear, Dropout)."""

        def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                     act_cfg=dict(type="ReLU", inplace=True), ffn_drop=0.0,
                     dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
            super().__init__(init_cfg)
            layers = []
            in_channels = embed_dims
            for _ in range(num_fcs - 1):
                layers.append(nn.Sequential(nn.Linear(in_channels, feedforward_channels),
                                            _build_activation_layer(act_cfg), nn.Dropout(ffn_drop)))
                in_channels = feedforward_channels
            layers.append(nn.Linear(feedforward_channels, embed_dims))
            layers.append(nn.Dropout(ffn_drop))
            self.layers = nn.Sequential(*layers)
            self.dropout_layer = _build_dropout(dropout_layer) if dropout_layer else nn.Identity()
            self.add_identity = add_identity
            self.embed_dims = embed_dims

        def forward(self, x, identity=None):
            out = self.layers(x)
            if not self.add_identity:
                return self.dropout_layer(out)
            if identity is None:
                identity = x
            return identity + self.dropout_layer(out)

    R["FEEDFORWARD_NETWORK"].register_module()(FFN)

    class MultiheadAttention(_BaseModule):
        """mmcv MultiheadAttention; the reference vendors a verbatim copy at
        petr_transformer.py:228-367, which this stub follows."""

        def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                     dropout_layer=dict(type="Dropout", drop_prob=0.), init_cfg=None,
                     batch_first=False, **kwargs):
            super().__init__(init_cfg)
            dropout_layer = dict(dropout_layer) if dropout_layer else dropout_layer
            if "dropout" in kwargs:
                attn_drop = kwargs["dropout"]
                dropout_layer["drop_prob"] = kwargs.pop("dropout")
            self.embed_dims = embed_dims
            self.num_heads = num_heads
            self.batch_first = batch_first
            self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
            self.proj_drop = nn.Dropout(proj_drop)
            self.dropout_layer = _build_dropout(dropout_layer) if dropout_layer else nn.Identity()

        def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                    key_pos=None, attn_mask=None, key_padding_mask=None, **kwargs):
            if key is None:
                key = query
            if value is None:
                value = key
            if identity is None:
                identity = query
            if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
                key_pos = query_pos
            if query_pos is not None:
                query = query + query_pos
            if key_pos is not None:
                key = key + key_pos
            if self.batch_first:
                query, key, value = (t.transpose(0, 1) for t in (query, key, value))
            out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                            key_padding_mask=key_padding_mask)[0]
            if self.batch_first:
                out = out.transpose(0, 1)
            return identity + self.dropout_layer(self.proj_drop(out))

    R["ATTENTION"].register_module()(MultiheadAttention)

    class BaseTransformerLayer(_BaseModule):
        """mmcv BaseTransformerLayer; follows the reference's vendored near-copy
        multi_atten_decoder_layer.py:72-293 (ctor :72-158, forward :160-293)."""

        def __init__(self, attn_cfgs=None,
                     ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=1024,
                                   num_fcs=2, ffn_drop=0., act_cfg=dict(type="ReLU", inplace=True)),
                     operation_order=None, norm_cfg=dict(type="LN"), init_cfg=None,
                     batch_first=False, **kwargs):
            deprecated = dict(feedforward_channels="feedforward_channels",
                              ffn_dropout="ffn_drop", ffn_num_fcs="num_fcs")
            ffn_cfgs = copy.deepcopy(ffn_cfgs)
            for ori, new in deprecated.items():
                if ori in kwargs:
                    ffn_cfgs[new] = kwargs[ori]
            super().__init__(init_cfg)
            self.batch_first = batch_first
            num_attn = operation_order.count("self_attn") + operation_order.count("cross_attn")
            if isinstance(attn_cfgs, dict):
                attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
            assert num_attn == len(attn_cfgs)
            self.num_attn = num_attn
            self.operation_order = operation_order
            self.norm_cfg = norm_cfg
            self.pre_norm = operation_order[0] == "norm"
            self.attentions = _ModuleList()
            index = 0
            for op in operation_order:
                if op in ("self_attn", "cross_attn"):
                    cfg = copy.deepcopy(attn_cfgs[index])
                    if "batch_first" in cfg:
                        assert self.batch_first == cfg["batch_first"]
                    else:
                        cfg["batch_first"] = self.batch_first
                    attention = R["ATTENTION"].build(cfg)
                    attention.operation_name = op
                    self.attentions.append(attention)
                    index += 1
            self.embed_dims = self.attentions[0].embed_dims
            self.ffns = _ModuleList()
            num_ffns = operation_order.count("ffn")
            if isinstance(ffn_cfgs, dict):
                ffn_cfgs = [copy.deepcopy(ffn_cfgs) for _ in range(num_ffns)]
            for i in range(num_ffns):
                if "embed_dims" not in ffn_cfgs[i]:
                    ffn_cfgs[i]["embed_dims"] = self.embed_dims
                else:
                    assert ffn_cfgs[i]["embed_dims"] == self.embed_dims
                self.ffns.append(R["FEEDFORWARD_NETWORK"].build(ffn_cfgs[i]))
            self.norms = _ModuleList()
            for _ in range(operation_order.count("norm")):
                self.norms.append(_build_norm_layer(norm_cfg, self.embed_dims)[1])

        def forward(self, query, key=None, value=None, query_pos=None, key_pos=None,
                    attn_masks=None, query_key_padding_mask=None, key_padding_mask=None,
                    **kwargs):
            norm_index = attn_index = ffn_index = 0
            identity = query
            if attn_masks is None:
                attn_masks = [None for _ in range(self.num_attn)]
            elif isinstance(attn_masks, torch.Tensor):
                attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
            for layer in self.operation_order:
                if layer == "self_attn":
                    temp_key = temp_value = query
                    query = self.attentions[attn_index](
                        query, temp_key, temp_value, identity if self.pre_norm else None,
                        query_pos=query_pos, key_pos=query_pos,
                        attn_mask=attn_masks[attn_index],
                        key_padding_mask=query_key_padding_mask, **kwargs)
                    attn_index += 1
                    identity = query
                elif layer == "norm":
                    query = self.norms[norm_index](query)
                    norm_index += 1
                elif layer == "cross_attn":
                    query = self.attentions[attn_index](
                        query, key, value, identity if self.pre_norm else None,
                        query_pos=query_pos, key_pos=key_pos,
                        attn_mask=attn_masks[attn_index],
                        key_padding_mask=key_padding_mask, **kwargs)
                    attn_index += 1
                    identity = query
                elif layer == "ffn":
                    query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                    ffn_index += 1
            return query

    R["TRANSFORMER_LAYER"].register_module()(BaseTransformerLayer)

    class TransformerLayerSequence(_BaseModule):
        def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
            super().__init__(init_cfg)
            if isinstance(transformerlayers, dict):
                transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
            assert len(transformerlayers) == num_layers
            self.num_layers = num_layers
            self.layers = _ModuleList()
            for i in range(num_layers):
                self.layers.append(R["TRANSFORMER_LAYER"].build(transformerlayers[i]))
            self.embed_dims = self.layers[0].embed_dims
            self.pre_norm = self.layers[0].pre_norm

        def forward(self, query, key, value, query_pos=None, key_pos=None, attn_masks=None,
                    query_key_padding_mask=None, key_padding_mask=None, **kwargs):
            for layer in self.layers:
                query = layer(query, key, value, query_pos=query_pos, key_pos=key_pos,
                              attn_masks=attn_masks,
                              query_key_padding_mask=query_key_padding_mask,
                              key_padding_mask=key_padding_mask, **kwargs)
            return query

    tfm.FFN = FFN
    tfm.MultiheadAttention = MultiheadAttention
    tfm.BaseTransformerLayer = BaseTransformerLayer
    tfm.TransformerLayerSequence = TransformerLayerSequence
    tfm.POSITIONAL_ENCODING = R["POSITIONAL_ENCODING"]
    tfm.build_positional_encoding = R["POSITIONAL_ENCODING"].build
    tfm.build_transformer_layer_sequence = R["TRANSFORMER_LAYER_SEQUENCE"].build
    tfm.build_attention = R["ATTENTION"].build
    tfm.build_feedforward_network = R["FEEDFORWARD_NETWORK"].build

    # ---- mmdet ----------------------------------------------------------------------
    mmdet = mod("mmdet")
    core = mod("mmdet.core")
    bbox = mod("mmdet.core.bbox")
    bbox_builder = mod("mmdet.core.bbox.builder")
    models = mod("mmdet.models")
    mutils = mod("mmdet.models.utils")
    mbuilder = mod("mmdet.models.utils.builder")
    mtransformer = mod("mmdet.models.utils.transformer")
    dense_heads = mod("mmdet.models.dense_heads")
    afh = mod("mmdet.models.dense_heads.anchor_free_head")

    def _unavailable(*a, **k):
        raise NotImplementedError("training-only mmdet helper; not part of the oracle")

    for name in ("bbox_cxcywh_to_xyxy", "bbox_xyxy_to_cxcywh", "build_assigner",
                 "build_sampler", "multi_apply", "reduce_mean"):
        setattr(core, name, _unavailable)

    class BaseBBoxCoder:
        def __init__(self, **kwargs):
            pass

    bbox.BaseBBoxCoder = BaseBBoxCoder
    bbox_builder.BBOX_CODERS = R["BBOX_CODERS"]
    bbox_builder.build_bbox_coder = R["BBOX_CODERS"].build

    class _Loss(nn.Module):
        def __init__(self, use_sigmoid=False, **kwargs):
            super().__init__()
            self.use_sigmoid = use_sigmoid

    def build_loss(cfg):
        return _Loss(use_sigmoid=cfg.get("use_sigmoid", False))

    class AnchorFreeHead(_BaseModule):
        """forward-only stand-in: the real mmdet class also builds default losses and calls
        _init_layers() once (changes only the RNG stream)."""

        def __init__(self, num_classes, in_channels, init_cfg=None, **kwargs):
            super().__init__(init_cfg)

    class NormedLinear(nn.Linear):
        """mmdet 2.24.1 normed_predictor.NormedLinear (published semantics; attribute spelling as upstream)."""

        def __init__(self, *args, tempearture=20, power=1.0, eps=1e-6, **kwargs):
            super().__init__(*args, **kwargs)
            self.tempearture, self.power, self.eps = tempearture, power, eps
            nn.init.normal_(self.weight, mean=0, std=0.01)
            if self.bias is not None:
                nn.init.constant_(self.bias, 0)

        def forward(self, x):
            weight_ = self.weight / (self.weight.norm(dim=1, keepdim=True).pow(self.power) + self.eps)
            x_ = x / (x.norm(dim=1, keepdim=True).pow(self.power) + self.eps)
            return F.linear(x_ * self.tempearture, weight_, self.bias)

    models.HEADS = R["HEADS"]
    models.NECKS = R["NECKS"]
    models.build_loss = build_loss
    mutils.build_transformer = R["TRANSFORMER"].build
    mutils.NormedLinear = NormedLinear
    mbuilder.TRANSFORMER = R["TRANSFORMER"]
    mbuilder.build_transformer = R["TRANSFORMER"].build
    mtransformer.inverse_sigmoid = _inverse_sigmoid
    afh.AnchorFreeHead = AnchorFreeHead
    dense_heads.AnchorFreeHead = AnchorFreeHead

    # ---- mmdet3d ---------------------------------------------------------------------
    mod("mmdet3d")
    mod("mmdet3d.core")
    mod("mmdet3d.core.bbox")
    coders = mod("mmdet3d.core.bbox.coders")
    coders.build_bbox_coder = R["BBOX_CODERS"].build
    return R


def _load_by_path(modname, relpath):
    if modname in sys.modules:
        return sys.modules[modname]
    path = os.path.join(_PLUGIN, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    parent, _, child = modname.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(m)
    return m


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's own classes (PETRHead, PETRv2Head, ...)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise FileNotFoundError(f"reference not found under {REF_ROOT}")
    R = _install_stubs()
    for name in ("projects", "projects.mmdet3d_plugin", "projects.mmdet3d_plugin.core",
                 "projects.mmdet3d_plugin.core.bbox", "projects.mmdet3d_plugin.core.bbox.coders",
                 "projects.mmdet3d_plugin.models", "projects.mmdet3d_plugin.models.utils",
                 "projects.mmdet3d_plugin.models.dense_heads", "projects.mmdet3d_plugin.models.necks"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
            parent, _, child = name.rpartition(".")
            if parent:
                setattr(sys.modules[parent], child, m)
    p = "projects.mmdet3d_plugin."
    _load_by_path(p + "core.bbox.array_converter", "core/bbox/array_converter.py")
    _load_by_path(p + "core.bbox.util", "core/bbox/util.py")
    coder = _load_by_path(p + "core.bbox.coders.nms_free_coder", "core/bbox/coders/nms_free_coder.py")
    posenc = _load_by_path(p + "models.utils.positional_encoding", "models/utils/positional_encoding.py")
    tfm = _load_by_path(p + "models.utils.petr_transformer", "models/utils/petr_transformer.py")
    head = _load_by_path(p + "models.dense_heads.petr_head", "models/dense_heads/petr_head.py")
    headv2 = _load_by_path(p + "models.dense_heads.petrv2_head", "models/dense_heads/petrv2_head.py")
    cpfpn = _load_by_path(p + "models.necks.cp_fpn", "models/necks/cp_fpn.py")
    ns = types.SimpleNamespace(
        registries=R, PETRHead=head.PETRHead, PETRv2Head=headv2.PETRv2Head,
        PETRTransformer=tfm.PETRTransformer, PETRMultiheadAttention=tfm.PETRMultiheadAttention,
        PETRTransformerDecoder=tfm.PETRTransformerDecoder,
        PETRTransformerDecoderLayer=tfm.PETRTransformerDecoderLayer,
        SinePositionalEncoding3D=posenc.SinePositionalEncoding3D,
        LearnedPositionalEncoding3D=posenc.LearnedPositionalEncoding3D,
        NMSFreeCoder=coder.NMSFreeCoder, pos2posemb3d=head.pos2posemb3d, CPFPN=cpfpn.CPFPN,
        inverse_sigmoid=_inverse_sigmoid, head_module=head, headv2_module=headv2)
    _LOADED = ns
    return ns


def read_head_cfg(config_relpath: str) -> dict:
    """exec()s a reference config file (plain python dict assignments) and returns a copy of
    ``model['pts_bbox_head']`` (e.g. 'projects/configs/petr/petr_vovnet_gridmask_p4_800x320.py')."""
    path = os.path.join(REF_ROOT, config_relpath)
    g = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), g)
    return copy.deepcopy(g["model"]["pts_bbox_head"])


def build_reference_head(config_relpath: str, seed: int = 0, **overrides):
    """Builds the reference head from its own config dict, fp32, eval mode, init_weights()."""
    ns = load_reference()
    cfg = read_head_cfg(config_relpath)
    cfg.update(overrides)
    typ = cfg.pop("type")
    cfg["train_cfg"] = None
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        head = getattr(ns, typ)(**cfg)
        head.init_weights()
    return head.eval()
