"""``backend="reference"``: the head composed from stock ``torch.nn`` calls on the model's own parameter containers.

The fused autograd function is once-differentiable; Grad-CAM++-style tooling differentiates twice through the head
(``torch.autograd.grad(..., create_graph=True)`` - src/services/XAI/models/cam.py:38-43,
interpretability/gradcam_plusplus.py).  For those callers ``MultimodalModel(..., backend="reference")`` (or
``model.backend = "reference"`` on an existing model: same parameters, same state dict) evaluates the same fusion
strings with PyTorch's own kernels, so every order of derivative exists.  It is an explicit opt-in for analysis code,
never selected automatically, never timed by bench.py, and not a fallback for a missing CUDA extension (importing
fusion_b200 without the built library still raises).

Semantics follow multimodalIntraInterModal.py:172-416 (one branch per fusion string), gatedResidualBlock.py:12-17 and
metablock.py:22-32; the always-executed-but-unused attentions of :193-197 are skipped (they influence neither the logits
nor any gradient).
"""
from __future__ import annotations

import torch

RG = "att-intramodal+residual+cross-attention-metadados"


def _attend(mha, q, kv):
    return mha(q, kv, kv)[0]


def _residual(block, q, kv):
    a = block.dropout(_attend(block.attn, q, kv))
    g = torch.sigmoid(block.gate_linear(q))
    return block.norm(g * a + (1 - g) * q)


def _metablock(mb, v, u):
    return torch.sigmoid(torch.tanh(v * mb.fb(u)) + mb.gb(u))


def reference_forward(model, img_feat, text_in):
    """logits[B, C] for features ``img_feat`` [B, F] and metadata / text features ``text_in``."""
    m = model.attention_mecanism
    if m == "no-metadata-without-mlp":
        return model.fc_visual_only(img_feat)
    txt_feat = model.text_fc(text_in) if model.text_fc is not None else text_in
    if m == "metablock":
        return model.fc_mlp_module_after_metablock_fusion_module(_metablock(model.meta_block, img_feat, txt_feat))
    pi = model.image_projector(img_feat)
    if m == "no-metadata":
        return model.fc_fusion(pi)
    pt = model.text_projector(txt_feat)
    iq, tq = pi.unsqueeze(0), pt.unsqueeze(0)                      # (1, B, D): sequence length 1 (:187-191)
    fuse = lambda a, b: model.fc_fusion(torch.cat([a.reshape(pi.shape), b.reshape(pt.shape)], dim=1))
    if m == "concatenation":
        return fuse(pi, pt)
    if m == "weighted":
        return fuse(torch.sigmoid(model.img_gate(pi)) * pi, torch.sigmoid(model.txt_gate(pt)) * pt)
    if m == "rg-att2fusefeatures":
        return model.fc_fusion_proj_feat2output(_residual(model.image_residual, tq, iq).squeeze(0))
    if m == "rg-att":
        return fuse(_residual(model.image_residual, iq, tq), _residual(model.text_residual, tq, iq))
    if m == "cross-attention-only":
        return fuse(_attend(model.image_cross_attention, iq, tq), _attend(model.text_cross_attention, tq, iq))
    if m == "residual+cross-attention-metadados":
        ir, tr = _residual(model.image_residual, iq, iq), _residual(model.text_residual, tq, tq)
        return fuse(_attend(model.image_cross_attention, ir, tr), _attend(model.text_cross_attention, tr, ir))
    ia, ta = _attend(model.image_self_attention, iq, iq), _attend(model.text_self_attention, tq, tq)
    if m == "att-intramodal":
        return fuse(ia, ta)
    if m == "att-intramodal+residual":
        return fuse(_residual(model.image_residual, iq, ia), _residual(model.text_residual, tq, ta))
    if m in ("crossattention", "gfcam", "cross-weights-after-crossattention"):
        ic = _attend(model.image_cross_attention, ia, ta).squeeze(0)
        tc = _attend(model.text_cross_attention, ta, ia).squeeze(0)
        if m == "crossattention":
            return fuse(ic, tc)
        gi, gt = torch.sigmoid(model.img_gate(ic)), torch.sigmoid(model.txt_gate(tc))
        return fuse(gi * ic, gt * tc) if m == "gfcam" else fuse(gt * ic, gi * tc)       # :231-235 swaps the gates
    if m.startswith(RG):
        ir, tr = _residual(model.image_residual, iq, ia), _residual(model.text_residual, tq, ta)
        ic, tc = _attend(model.image_cross_attention, ir, tr), _attend(model.text_cross_attention, tr, ir)
        tail = m[len(RG):]
        if tail == "":
            return fuse(ic, tc)
        if tail == "+rg-att2fusefeatures":
            return model.fc_fusion_proj_feat2output(_residual(model.image_residual, tc, ic).squeeze(0))
        if tail == "+metablock":
            return model.fc_fusion_proj_feat2output(_metablock(model.meta_block, ic.squeeze(0), tc.squeeze(0)))
        if tail == "+att-intramodal+residual":
            ia2, ta2 = _attend(model.image_self_attention, ic, ic), _attend(model.text_self_attention, tc, tc)
            return fuse(_residual(model.image_residual, ic, ia2), _residual(model.text_residual, tc, ta2))
    raise ValueError(f"Attention mechanism '{m}' not implemented.")
