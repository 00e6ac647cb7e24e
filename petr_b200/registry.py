"""Registries with the mmcv ``Registry`` surface the reference configs rely on
(``type='PETRHead'`` string dispatch; petr_head.py:46, petr_transformer.py:33,:112,:227,:400,
positional_encoding.py:14).  When mmcv/mmdet are importable the same classes are also
registered there (``force=True``) so ``plugin=True`` configs pick up this implementation;
otherwise these stand-alone registries are the build path.
"""
from __future__ import annotations

import copy


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise KeyError(f"cfg for {self.name} must be a dict with a 'type' key, got {cfg!r}")
        cfg = copy.deepcopy(dict(cfg))
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        typ = cfg.pop("type")
        cls = self.module_dict.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**cfg)


ATTENTION = Registry("attention")
FEEDFORWARD_NETWORK = Registry("feed-forward network")
TRANSFORMER_LAYER = Registry("transformerLayer")
TRANSFORMER_LAYER_SEQUENCE = Registry("transformer-layers sequence")
POSITIONAL_ENCODING = Registry("position encoding")
TRANSFORMER = Registry("Transformer")
HEADS = Registry("head")
BBOX_CODERS = Registry("bbox_coder")
NECKS = Registry("neck")

build_attention = ATTENTION.build
build_feedforward_network = FEEDFORWARD_NETWORK.build
build_transformer_layer = TRANSFORMER_LAYER.build
build_transformer_layer_sequence = TRANSFORMER_LAYER_SEQUENCE.build
build_positional_encoding = POSITIONAL_ENCODING.build
build_transformer = TRANSFORMER.build
build_head = HEADS.build
build_bbox_coder = BBOX_CODERS.build
build_neck = NECKS.build


def register_into_mm(verbose=False) -> bool:
    """Registers every class above into the real mmcv / mmdet registries if present."""
    try:
        from mmcv.cnn.bricks import registry as mmcv_reg          # type: ignore
        from mmdet.models.utils.builder import TRANSFORMER as MM_TRANSFORMER   # type: ignore
        from mmdet.models import HEADS as MM_HEADS                # type: ignore
        from mmdet.core.bbox.builder import BBOX_CODERS as MM_CODERS  # type: ignore
        from mmdet.models import NECKS as MM_NECKS                # type: ignore
    except Exception:
        return False
    pairs = [(ATTENTION, mmcv_reg.ATTENTION), (TRANSFORMER_LAYER, mmcv_reg.TRANSFORMER_LAYER),
             (TRANSFORMER_LAYER_SEQUENCE, mmcv_reg.TRANSFORMER_LAYER_SEQUENCE),
             (POSITIONAL_ENCODING, mmcv_reg.POSITIONAL_ENCODING),
             (TRANSFORMER, MM_TRANSFORMER), (HEADS, MM_HEADS), (BBOX_CODERS, MM_CODERS),
             (NECKS, MM_NECKS)]
    for ours, theirs in pairs:
        for name, cls in ours.module_dict.items():
            if name in ("MultiheadAttention", "FFN"):
                continue                      # leave mmcv's own generic bricks alone
            theirs.register_module(name=name, force=True, module=cls)
            if verbose:
                print(f"petr_b200: registered {name} into {theirs.name}")
    return True
