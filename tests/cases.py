"""Shared parity cases: seeded inputs + weights for the golden fixtures, the oracle tests and
the GPU parity tests.  A case is fully determined by its entry below (plus petr_b200.synth)."""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from petr_b200 import synth
from petr_b200.configs import WORKLOADS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    # name: workload, batch, num_query, seed, extras
    "tiny": dict(workload="tiny", batch=2, num_query=64, seed=0),
    "tiny_padded": dict(workload="tiny", batch=3, num_query=64, seed=1,
                        img_shape=(40, 130), pad_shape=(64, 160), fully_masked=[1]),
    "tiny_sharp": dict(workload="tiny", batch=2, num_query=64, seed=2, qk_scale=8.0, feat_scale=4.0),
    "tiny_range": dict(workload="tiny", batch=1, num_query=64, seed=3,
                       position_range=[-20.0, -20.0, -2.0, 20.0, 20.0, 2.0]),
    "tiny_v2": dict(workload="tiny-v2", batch=1, num_query=64, seed=4),
    "vov_800x320": dict(workload="vov-p4-800x320", batch=1, num_query=900, seed=0),
    # 1500-key cases for the bf16 (tcgen05) path: bf16 rounding of V / P / position embeddings
    # averages out as 1/sqrt(keys); at the 80 keys of the tiny cases it does not
    "small": dict(workload="small", batch=2, num_query=64, seed=5),
    "small_padded": dict(workload="small", batch=3, num_query=64, seed=6,
                         img_shape=(140, 390), pad_shape=(160, 400), fully_masked=[1]),
    "small_sharp": dict(workload="small", batch=2, num_query=64, seed=7, qk_scale=3.0),
    "small_v2": dict(workload="small-v2", batch=1, num_query=64, seed=8),
}

TINY = [n for n in CASES if n.startswith("tiny")]
BF16_CASES = ["small", "small_padded", "small_sharp", "small_v2", "vov_800x320"]


def head_cfg(case: dict) -> dict:
    wl = WORKLOADS[case["workload"]]
    cfg = wl.head_cfg()
    cfg["num_query"] = case["num_query"]
    if "position_range" in case:
        cfg["position_range"] = list(case["position_range"])
    return cfg


def oracle_cfg(case: dict) -> dict:
    wl = WORKLOADS[case["workload"]]
    v2 = wl.head == "PETRv2Head"
    c = dict(num_query=case["num_query"], with_fpe=v2, with_time=v2, with_multi=v2)
    if "position_range" in case:
        c["position_range"] = tuple(case["position_range"])
    return c


def inputs(case: dict):
    """Returns (workload, feats cpu fp32 (B,N,C,H,W), img_metas)."""
    wl = WORKLOADS[case["workload"]]
    B, seed = case["batch"], case["seed"]
    metas = synth.make_img_metas(wl, B, seed, case.get("img_shape"), case.get("pad_shape"))
    for b in case.get("fully_masked", []):
        metas[b]["img_shape"] = [(0, 0, 3)] * wl.num_cams
    feats = synth.make_features(wl, B, seed, scale=case.get("feat_scale", 1.0))
    return wl, feats, metas


def weights(case: dict, head) -> dict:
    sd = synth.head_state_dict(head, case["seed"])
    if "qk_scale" in case:
        sd = synth.sharpen_state_dict(sd, case["qk_scale"])
    return sd


def golden_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, f"{name}.npz")


POS_SAMPLE_STRIDE = 37     # large cases store every 37th element of coords_position_embeding


def load_golden(name: str) -> dict:
    with np.load(golden_path(name)) as z:
        return {k: z[k] for k in z.files}
