"""Deterministic case definitions shared by the golden generator and the tests.

Weights, inputs, labels and dropout keep-masks are drawn from numpy's PCG64 (stable
across numpy versions and machines), NOT from torch's RNG, so the GPU box can rebuild
the exact inputs of every fixture without the reference or the fixture storing them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.head_oracle import HeadConfig, RG_ATT, IN_SCOPE, MECHANISMS  # noqa: E402

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))

# name -> dict(cfg kwargs, B, seed, train)
# "full" cases = BASELINE.json configs; "small" cases carry full gradients.
FULL = {
    "cfg1_concat":  dict(mechanism="concatenation", F=512, V=85, C=6),
    "cfg2_cross":   dict(mechanism="crossattention", F=2048, V=85, C=6),
    "cfg3a_meta":   dict(mechanism="metablock", F=1664, V=13, C=8),
    "cfg3b_weight": dict(mechanism="weighted", F=1664, V=13, C=8),
    "cfg4a_gfcam":  dict(mechanism="gfcam", F=768, V=None, T=85, C=2, text_model="tab-transformer"),
    "cfg4b_rgatt":  dict(mechanism=RG_ATT, F=768, V=None, T=85, C=2, text_model="tab-transformer"),
    "cfg5_rgatt":   dict(mechanism=RG_ATT, F=1024, V=85, C=6),
}
SMALL_DIMS = dict(F=48, V=13, C=5, T=40, D=64, H=8)


def all_cases():
    cases = {}
    for name, kw in FULL.items():
        for train in (False, True):
            cases[f"{name}_{'train' if train else 'eval'}"] = dict(cfg=kw, B=32, seed=101 + len(cases), train=train, full_grads=False)
    for i, mech in enumerate(MECHANISMS):
        kw = dict(SMALL_DIMS, mechanism=mech)
        tag = f"small{i:02d}"
        cases[f"{tag}_train"] = dict(cfg=kw, B=5, seed=900 + i, train=True, full_grads=True)
    # ragged / edge batches on the headline mechanism
    for B in (1, 33):
        cases[f"edge_cross_B{B}"] = dict(cfg=dict(SMALL_DIMS, mechanism="crossattention"), B=B, seed=700 + B, train=False, full_grads=True)
    return cases


def make_cfg(kw) -> HeadConfig:
    return HeadConfig(**kw)


def mask_keys(cfg: HeadConfig):
    """Dropout sites in reference call order (nn.Dropout modules reached by forward)."""
    m = cfg.mechanism
    keys = []
    uses_res = {
        "rg-att2fusefeatures": ["img_res"],
        "rg-att": ["img_res", "txt_res"],
        "att-intramodal+residual": ["img_res", "txt_res"],
        "residual+cross-attention-metadados": ["img_res", "txt_res"],
        RG_ATT: ["img_res", "txt_res"],
        RG_ATT + "+rg-att2fusefeatures": ["img_res", "txt_res", "img_res2"],
        RG_ATT + "+metablock": ["img_res", "txt_res"],
        RG_ATT + "+att-intramodal+residual": ["img_res", "txt_res", "img_res2", "txt_res2"],
    }
    keys += uses_res.get(m, [])
    no_mlp = {"no-metadata-without-mlp", "rg-att2fusefeatures", RG_ATT + "+rg-att2fusefeatures", RG_ATT + "+metablock"}
    if m not in no_mlp:
        keys += ["fc1", "fc2"]
    return keys


def mask_shape(cfg: HeadConfig, key, B):
    if key == "fc2":
        return (B, cfg.D // 2)
    return (B, cfg.D)


def mask_p(cfg: HeadConfig, key):
    if key.startswith(("img_res", "txt_res")):
        return 0.1
    return 0.3 if cfg.mechanism == "metablock" else 0.5


_LN_MARKS = (".norm.", "fc_fusion.1.", "fc_fusion.5.", "meta_block.fb.1.", "meta_block.gb.1.",
             "fusion_module.1.", "fusion_module.5.")


def is_ln(name):
    return any(k in name for k in _LN_MARKS)


def gen_params(cfg: HeadConfig, seed, dtype=np.float32):
    """Random head parameters: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) matrices, non-zero biases
    (also for the MHA in_proj biases, which torch zero-initialises), LN gamma ~ 1+-0.2."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for name, shp in cfg.param_shapes().items():
        if len(shp) == 2:
            bound = 1.0 / np.sqrt(shp[1])
            a = rng.uniform(-bound, bound, size=shp)
        elif is_ln(name):
            a = 1.0 + 0.2 * rng.standard_normal(shp) if name.endswith("weight") else 0.1 * rng.standard_normal(shp)
        else:
            a = 0.1 * rng.standard_normal(shp)
        out[name] = a.astype(dtype)
    return out


def gen_inputs(cfg: HeadConfig, B, seed, train, dtype=np.float32):
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    x = rng.standard_normal((B, cfg.F)).astype(dtype)
    tw = cfg.V if cfg.text_model == "one-hot-encoder" else cfg.T
    tin = rng.standard_normal((B, tw)).astype(dtype)
    labels = rng.integers(0, cfg.C, size=(B,)).astype(np.int64)
    counts = np.bincount(labels, minlength=cfg.C).astype(np.float64)
    # class weights N/(C*count_c) (train_pad_20.py:22-32); unseen classes get weight 1
    cw = np.where(counts > 0, B / (cfg.C * np.maximum(counts, 1)), 1.0).astype(dtype)
    masks = None
    if train:
        masks = {}
        for k in mask_keys(cfg):
            masks[k] = (rng.random(mask_shape(cfg, k, B)) >= mask_p(cfg, k)).astype(np.uint8)
    return x, tin, labels, cw, masks


# -- compact gradient summaries for the full-size cases --------------------------------
FULL_GRAD_MAX_ELEMS = 20000
N_PROBES = 4
N_SAMPLES = 64


def grad_summary(name, g, seed):
    """Size-independent pin of one gradient tensor: a few seeded random projections,
    the max-abs, and N_SAMPLES entries at seeded flat indices."""
    rng = np.random.Generator(np.random.PCG64(abs(hash_name(name)) + seed))
    flat = np.asarray(g, dtype=np.float64).ravel()
    probes = np.array([float(flat @ rng.standard_normal(flat.size)) for _ in range(N_PROBES)])
    idx = rng.integers(0, flat.size, size=N_SAMPLES)
    return dict(probes=probes, idx=idx, samples=flat[idx].copy(), maxabs=float(np.abs(flat).max()),
                l2=float(np.sqrt((flat * flat).sum())))


def hash_name(name):
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h
