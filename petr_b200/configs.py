"""Head config dicts and named workloads.

``petr_head_cfg`` / ``petrv2_head_cfg`` build the same ``model['pts_bbox_head']`` dicts the
reference's config files hold (projects/configs/petr/petr_vovnet_gridmask_p4_1600x640.py:43-96,
petr_r50dcn_gridmask_c5.py:45-98, projects/configs/petrv2/petrv2_vovnet_gridmask_p4_800x320.py:41-101)
so the GPU box can build heads without /root/reference.  ``tests/test_configs.py`` checks them
key-for-key against the reference files when those are present.
"""
from __future__ import annotations

from dataclasses import dataclass

POINT_CLOUD_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
VOXEL_SIZE = [0.2, 0.2, 8]
POSITION_RANGE = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]


def _transformer_cfg(num_layers=6, embed_dims=256, num_heads=8, feedforward_channels=2048,
                     with_cp=True):
    cfg = dict(
        type='PETRTransformer',
        decoder=dict(
            type='PETRTransformerDecoder',
            return_intermediate=True,
            num_layers=num_layers,
            transformerlayers=dict(
                type='PETRTransformerDecoderLayer',
                attn_cfgs=[
                    dict(type='MultiheadAttention', embed_dims=embed_dims,
                         num_heads=num_heads, dropout=0.1),
                    dict(type='PETRMultiheadAttention', embed_dims=embed_dims,
                         num_heads=num_heads, dropout=0.1),
                ],
                feedforward_channels=feedforward_channels,
                ffn_dropout=0.1,
                with_cp=True,
                operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')),
        ))
    if with_cp is None:          # petr_vovnet_gridmask_p4_800x320.py leaves with_cp at its default
        del cfg['decoder']['transformerlayers']['with_cp']
    return cfg


def _common(in_channels, with_cp=True):
    return dict(
        num_classes=10,
        in_channels=in_channels,
        num_query=900,
        LID=True,
        with_position=True,
        with_multiview=True,
        position_range=list(POSITION_RANGE),
        transformer=_transformer_cfg(with_cp=with_cp),
        bbox_coder=dict(
            type='NMSFreeCoder',
            post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
            pc_range=list(POINT_CLOUD_RANGE),
            max_num=300,
            voxel_size=list(VOXEL_SIZE),
            num_classes=10),
        positional_encoding=dict(type='SinePositionalEncoding3D', num_feats=128, normalize=True),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
        loss_bbox=dict(type='L1Loss', loss_weight=0.25),
        loss_iou=dict(type='GIoULoss', loss_weight=0.0))


def petr_head_cfg(in_channels=256, with_cp=True, **overrides):
    cfg = dict(type='PETRHead', normedlinear=False)
    cfg.update(_common(in_channels, with_cp))
    cfg.update(overrides)
    return cfg


def petrv2_head_cfg(in_channels=256, **overrides):
    cfg = dict(type='PETRv2Head', with_fpe=True, with_time=True, with_multi=True,
               code_weights=[1.0] * 10)
    cfg.update(_common(in_channels))
    cfg.update(overrides)
    return cfg


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json config: token shapes of the synthetic head input."""
    name: str
    head: str          # 'PETRHead' | 'PETRv2Head'
    num_cams: int
    in_channels: int
    feat_h: int
    feat_w: int
    img_h: int
    img_w: int

    @property
    def num_keys(self):
        return self.num_cams * self.feat_h * self.feat_w

    def head_cfg(self):
        return petrv2_head_cfg(self.in_channels) if self.head == 'PETRv2Head' \
            else petr_head_cfg(self.in_channels)


WORKLOADS = {
    w.name: w for w in [
        Workload('vov-p4-800x320', 'PETRHead', 6, 256, 20, 50, 320, 800),
        Workload('r50-c5-1408x512', 'PETRHead', 6, 2048, 16, 44, 512, 1408),
        Workload('vov-p4-1600x640', 'PETRHead', 6, 256, 40, 100, 640, 1600),
        Workload('v2-vov-p4-800x320', 'PETRv2Head', 12, 256, 20, 50, 320, 800),
        # small shapes for fast tests (not BASELINE configs)
        Workload('tiny', 'PETRHead', 2, 64, 4, 10, 64, 160),
        Workload('tiny-v2', 'PETRv2Head', 12, 64, 4, 10, 64, 160),
        # 1500-key shapes: small enough for the CPU oracle, enough keys for bf16 statistics
        Workload('small', 'PETRHead', 6, 64, 10, 25, 160, 400),
        Workload('small-v2', 'PETRv2Head', 12, 64, 5, 25, 80, 400),
    ]
}
