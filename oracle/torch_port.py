"""CPU timing port of the reference model (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference is pure Python on top of torch.nn, and /root/reference does not exist on the
GPU box, so the `cpu_baseline` / `--impl reference` legs of bench.py time THIS port: the
same torch.nn modules, called in the same order and with the same dead work as the
reference's forward (all four nn.MultiheadAttention calls run for every fusion string,
multimodalIntraInterModal.py:193-197; image_projector / text_projector run even for
`metablock`), so its CPU cost is the reference's CPU cost.  tests/test_torch_port.py pins
it against the unmodified reference (bit-identical outputs under the same seed) whenever
/root/reference is present.  Nothing in the product package imports this file.
"""
from __future__ import annotations

import torch
import torch.nn as nn

RG = "att-intramodal+residual+cross-attention-metadados"


def _mlp(first_in, D, C, p):
    return nn.Sequential(nn.Linear(first_in, D), nn.LayerNorm(D), nn.ReLU(), nn.Dropout(p),
                         nn.Linear(D, D // 2), nn.LayerNorm(D // 2), nn.ReLU(), nn.Dropout(p), nn.Linear(D // 2, C))


class _Residual(nn.Module):                      # gatedResidualBlock.py:4-17
    def __init__(self, D):
        super().__init__()
        self.norm = nn.LayerNorm(D)
        self.attn = nn.MultiheadAttention(embed_dim=D, num_heads=8, batch_first=False)
        self.dropout = nn.Dropout(0.1)
        self.gate_linear = nn.Linear(D, D)

    def forward(self, q, k, v):
        a = self.dropout(self.attn(q, k, v)[0])
        g = torch.sigmoid(self.gate_linear(q))
        return self.norm(g * a + (1 - g) * q)


class _Meta(nn.Module):                          # metablock.py:4-32
    def __init__(self, vdim, udim):
        super().__init__()
        self.fb = nn.Sequential(nn.Linear(udim, vdim), nn.LayerNorm(vdim))
        self.gb = nn.Sequential(nn.Linear(udim, vdim), nn.LayerNorm(vdim))

    def forward(self, V, U):
        return torch.sigmoid(torch.tanh(V * self.fb(U)) + self.gb(U))


class ReferencePort(nn.Module):
    """Head of MultimodalModel with an identity backbone: forward(img_feat[B,F], text_in)."""

    def __init__(self, mechanism, F, C, V=85, T=512, D=512, H=8, n=2, one_hot=True):
        super().__init__()
        self.m, self.one_hot = mechanism, one_hot
        self.image_projector = nn.Linear(F, D)
        if one_hot:
            self.text_fc = nn.Sequential(nn.Linear(V, 256), nn.ReLU(), nn.Linear(256, 512), nn.ReLU(), nn.Linear(512, T))
        self.text_projector = nn.Linear(T, D)
        for k in ("image_self_attention", "text_self_attention", "image_cross_attention", "text_cross_attention"):
            setattr(self, k, nn.MultiheadAttention(embed_dim=D, num_heads=H, batch_first=False))
        self.img_gate, self.txt_gate = nn.Linear(D, D), nn.Linear(D, D)
        common = mechanism == RG + "+metablock"
        self.meta_block = _Meta(D if common else F, D if common else T)
        self.image_residual, self.text_residual = _Residual(D), _Residual(D)
        self.fc_fusion = _mlp(D * (1 if mechanism == "no-metadata" else n), D, C, 0.5)
        self.fc_visual_only = nn.Linear(F, C)
        self.fc_fusion_proj_feat2output = nn.Linear(D, C)
        self.fc_mlp_module_after_metablock_fusion_module = _mlp(F, D, C, 0.3)

    def forward(self, img_feat, text_in):
        m = self.m
        pi = self.image_projector(img_feat)
        tf = self.text_fc(text_in) if self.one_hot else text_in
        pt = self.text_projector(tf)
        iq, tq = pi.unsqueeze(0), pt.unsqueeze(0)
        ia = self.image_self_attention(iq, iq, iq)[0]
        ta = self.text_self_attention(tq, tq, tq)[0]
        ic = self.image_cross_attention(ia, ta, ta)[0]
        tc = self.text_cross_attention(ta, ia, ia)[0]
        ip, tp = ic.squeeze(0), tc.squeeze(0)
        fuse = lambda a, b: self.fc_fusion(torch.cat([a, b], dim=1))
        cross = lambda a, b: (self.image_cross_attention(a, b, b)[0], self.text_cross_attention(b, a, a)[0])
        if m == "no-metadata":
            return self.fc_fusion(pi)
        if m == "no-metadata-without-mlp":
            return self.fc_visual_only(img_feat)
        if m == "concatenation":
            return fuse(pi, pt)
        if m == "crossattention":
            return fuse(ip, tp)
        if m == "weighted":
            return fuse(torch.sigmoid(self.img_gate(pi)) * pi, torch.sigmoid(self.txt_gate(pt)) * pt)
        if m in ("gfcam", "cross-weights-after-crossattention"):
            gi, gt = torch.sigmoid(self.img_gate(ip)), torch.sigmoid(self.txt_gate(tp))
            return fuse(gi * ip, gt * tp) if m == "gfcam" else fuse(gt * ip, gi * tp)
        if m == "metablock":
            return self.fc_mlp_module_after_metablock_fusion_module(self.meta_block(img_feat, tf))
        if m == "rg-att2fusefeatures":
            return self.fc_fusion_proj_feat2output(self.image_residual(tq, iq, iq).squeeze(0))
        if m == "rg-att":
            return fuse(self.image_residual(iq, tq, tq).squeeze(0), self.text_residual(tq, iq, iq).squeeze(0))
        if m == "att-intramodal":
            return fuse(ia.squeeze(0), ta.squeeze(0))
        if m == "att-intramodal+residual":
            return fuse(self.image_residual(iq, ia, ia).squeeze(0), self.text_residual(tq, ta, ta).squeeze(0))
        if m == "cross-attention-only":
            a, b = cross(iq, tq)
            return fuse(a.squeeze(0), b.squeeze(0))
        if m == "residual+cross-attention-metadados":
            a, b = cross(self.image_residual(iq, iq, iq), self.text_residual(tq, tq, tq))
            return fuse(a.squeeze(0), b.squeeze(0))
        if m.startswith(RG):
            a, b = cross(self.image_residual(iq, ia, ia), self.text_residual(tq, ta, ta))
            tail = m[len(RG):]
            if tail == "":
                return fuse(a.squeeze(0), b.squeeze(0))
            if tail == "+rg-att2fusefeatures":
                return self.fc_fusion_proj_feat2output(self.image_residual(b, a, a).squeeze(0))
            if tail == "+metablock":
                return self.fc_fusion_proj_feat2output(self.meta_block(a.squeeze(0), b.squeeze(0)))
            if tail == "+att-intramodal+residual":
                a2 = self.image_self_attention(a, a, a)[0]
                b2 = self.text_self_attention(b, b, b)[0]
                return fuse(self.image_residual(a, a2, a2).squeeze(0), self.text_residual(b, b2, b2).squeeze(0))
        raise ValueError(f"Attention mechanism '{m}' not implemented.")


def time_cpu_train_step(mechanism, F, V, C, B, T=512, D=512, H=8, one_hot=True, steps=5, warmup=2, threads=None,
                        budget_s=20.0, seed=1234):
    """Reference train step on the host cores: zero_grad + forward + weighted CE + backward
    (BASELINE.md section 4).  Returns dict(samples_per_s, ms_per_step, steps, threads)."""
    import time
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    model = ReferencePort(mechanism, F, C, V=V, T=T, D=D, H=H, one_hot=one_hot).train()
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(B, F, generator=g)
    t = torch.randn(B, V if one_hot else T, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    counts = torch.bincount(y, minlength=C).clamp_min(1).float()
    crit = nn.CrossEntropyLoss(weight=B / (C * counts))
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        loss = crit(model(x, t), y)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    times.sort()
    med = times[len(times) // 2]
    return dict(samples_per_s=B / med, ms_per_step=med * 1e3, steps=len(times), threads=torch.get_num_threads())
