"""CPU: the drop-in boundary -- registry/config construction, checkpoint (state_dict) contract,
legacy-key remap, error behaviour without a GPU, and that libpetr_b200.so loads and exports every
symbol include/petr_b200.h declares."""
import ctypes
import json
import os
import re

import pytest
import torch

import petr_b200 as pb
from petr_b200 import _lib
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _keys():
    with open(os.path.join(cases.GOLDEN_DIR, "state_dict_keys.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("label,cfg", [("PETRHead", pb.petr_head_cfg(256)),
                                       ("PETRHead_c5", pb.petr_head_cfg(2048)),
                                       ("PETRv2Head", pb.petrv2_head_cfg(256))])
def test_state_dict_layout_matches_reference(label, cfg):
    head = pb.build_head(cfg)
    mine = {k: list(v.shape) for k, v in head.state_dict().items()}
    assert mine == _keys()[label]


def test_shared_vs_copied_branches():
    v1 = pb.build_head(pb.petr_head_cfg(256))
    v2 = pb.build_head(pb.petrv2_head_cfg(256))
    assert v1.cls_branches[0] is v1.cls_branches[5] and v1.reg_branches[0] is v1.reg_branches[5]
    assert v2.cls_branches[0] is not v2.cls_branches[5]
    assert sum(p.numel() for p in v1.parameters()) == 11091882
    assert sum(p.numel() for p in v2.parameters()) == 14543886


def test_legacy_key_remap():
    """petr_head.py:346-359: version<2 checkpoints use .self_attn./.multihead_attn./.decoder.norm."""
    head = pb.build_head(pb.petr_head_cfg(256))
    sd = head.state_dict()
    legacy = {}
    for k, v in sd.items():
        k = k.replace(".attentions.0.", ".self_attn.").replace(".attentions.1.", ".multihead_attn.")
        k = k.replace(".decoder.post_norm.", ".decoder.norm.")
        legacy[k] = v.clone() + 1.0
    assert any(".self_attn." in k for k in legacy)
    head2 = pb.build_head(pb.petr_head_cfg(256))
    head2.load_state_dict(legacy, strict=True)
    for k, v in head2.state_dict().items():
        assert torch.equal(v, sd[k] + 1.0), k


def test_registry_and_ctor_kwargs():
    for name in ["PETRHead", "PETRv2Head"]:
        assert pb.HEADS.get(name) is not None
    for name in ["PETRMultiheadAttention", "MultiheadAttention"]:
        assert pb.ATTENTION.get(name) is not None
    assert pb.TRANSFORMER.get("PETRTransformer") is not None
    assert pb.TRANSFORMER_LAYER.get("PETRTransformerDecoderLayer") is not None
    for name in ["PETRTransformerDecoder", "PETRTransformerEncoder"]:
        assert pb.TRANSFORMER_LAYER_SEQUENCE.get(name) is not None
    for name in ["SinePositionalEncoding3D", "LearnedPositionalEncoding3D"]:
        assert pb.POSITIONAL_ENCODING.get(name) is not None
    head = pb.build_head(pb.petr_head_cfg(256))
    assert head.num_pred == 6 and head.embed_dims == 256 and head.position_level == 0
    assert head.position_dim == 192 and head.cls_out_channels == 10 and head.code_size == 10
    assert head.pc_range == [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    att = head.transformer.decoder.layers[0].attentions[1]
    assert isinstance(att, pb.PETRMultiheadAttention) and att.attn.dropout == 0.1
    with pytest.raises(KeyError):
        pb.build_head(dict(type="NoSuchHead"))


def test_no_cpu_fallback():
    head = pb.build_head(pb.petr_head_cfg(64, num_query=16)).eval()
    wl = pb.WORKLOADS["tiny"]
    from petr_b200 import synth
    x = synth.make_features(wl, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        head([x], synth.make_img_metas(wl, 1))
    head.train()                      # the training path is CUDA-only as well
    with pytest.raises(RuntimeError, match="CUDA"):
        head([x], synth.make_img_metas(wl, 1))
    att = head.transformer.decoder.layers[0].attentions[0]
    with pytest.raises(RuntimeError, match="inference"):       # stand-alone bricks: the head drives training
        att(x.new_zeros(4, 1, 256))


def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "petr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(petr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 12
    L = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in petr_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().petr_abi_version() == _lib.ABI_VERSION == 3


def test_argument_errors_are_reported_not_thrown(lib_built):
    L = _lib.lib()
    rc = L.petr_layernorm_f32(None, None, None, None, 4, 256, 1e-5, 0, 0, None)
    assert rc == -1
    assert b"null" in L.petr_last_error_string()
    rc = L.petr_key_padding_mask(None, 1, 1, 1, 1, 1, 1, None, None)
    assert rc < 0


def test_nms_free_coder_has_no_cpu_fallback():
    """NMSFreeCoder.decode / get_bboxes run one CUDA kernel (petr_nms_free_decode); CPU tensors are refused."""
    outs = dict(all_cls_scores=torch.randn(6, 2, 900, 10), all_bbox_preds=torch.randn(6, 2, 900, 10))
    head = pb.build_head(pb.petr_head_cfg(256))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head.bbox_coder.decode(outs)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head.get_bboxes(outs, [dict(), dict()])


def test_stacked_img_metas_ingestion_equals_list_format():
    """a pre-stacked dict of arrays is ingested identically to the list of dicts."""
    from petr_b200 import engine, synth
    for wn, kw in (("vov-p4-800x320", {}), ("v2-vov-p4-800x320", {}),
                   ("tiny", dict(img_shape=(40, 130), pad_shape=(64, 160)))):
        wl = pb.WORKLOADS[wn]
        metas = synth.make_img_metas(wl, 3, 4, **kw)
        if wn == "tiny":
            metas[1]["img_shape"] = [(0, 0, 3)] * wl.num_cams
        a = engine.host_prep(metas, wl.num_cams, wl.feat_h, wl.feat_w, wl.head == "PETRv2Head")
        b = engine.host_prep(engine.stack_img_metas(metas), wl.num_cams, wl.feat_h, wl.feat_w, wl.head == "PETRv2Head")
        assert a.keys() == b.keys()
        for k in a:
            if isinstance(a[k], torch.Tensor):
                assert torch.equal(a[k], b[k]), k
            else:
                assert a[k] == b[k], k
    assert a["has_mask"] and not a["uniform_masks"]


def test_engine_cache_follows_parameter_updates():
    """The packed-weight / CUDA-graph cache must not survive a checkpoint load that goes through a
    parent module's recursion (detector.load_state_dict, mmcv load_checkpoint: the head's own
    load_state_dict is never called) nor an in-place parameter update."""
    head = pb.build_head(pb.petr_head_cfg(256))
    eng = head.engine
    dev = torch.device("cpu")

    class Detector(torch.nn.Module):
        def __init__(self, h):
            super().__init__()
            self.pts_bbox_head = h

    det = Detector(head)
    eng._packed[dev] = {"stale": True}
    det.load_state_dict({k: v.clone() + 0.5 for k, v in det.state_dict().items()}, strict=True)
    assert head.engine is eng and dev not in eng._packed
    eng._packed[dev] = {"stale": True}
    with torch.no_grad():
        head.reference_points.weight.mul_(0.5)
    assert dev not in head.engine._packed
    eng._packed[dev] = {"fresh": True}
    assert dev in head.engine._packed              # nothing changed: the cache is kept
