"""Golden vectors for the auxiliary losses, from the UNMODIFIED reference classes (build container only):
    python tests/golden/make_golden_losses.py  ->  tests/golden/losses.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, "/root/reference/src/scripts/benchmark/models")
from focalLoss import FocalLoss  # noqa: E402
from softtargetsCrossEntropy import SoftTargetCrossEntropy  # noqa: E402

rng = np.random.Generator(np.random.PCG64(2024))
out = {}
for tag, (B, C) in {"a": (7, 6), "b": (33, 8), "c": (64, 2)}.items():
    z = (2 * rng.standard_normal((B, C)))
    y = rng.integers(0, C, B)
    alpha = rng.random(C) + 0.25
    t = rng.random((B, C)); t /= t.sum(1, keepdims=True)
    w = rng.random(C) + 0.5
    for name, use_w in (("w", True), ("n", False)):
        zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
        loss = FocalLoss(alpha=torch.tensor(alpha) if use_w else None, gamma=2)(zt, torch.tensor(y))
        loss.backward()
        out[f"focal_{tag}{name}_loss"], out[f"focal_{tag}{name}_dz"] = loss.item(), zt.grad.numpy().copy()
        zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
        loss = SoftTargetCrossEntropy(weight=torch.tensor(w) if use_w else None)(zt, torch.tensor(t))
        loss.backward()
        out[f"soft_{tag}{name}_loss"], out[f"soft_{tag}{name}_dz"] = loss.item(), zt.grad.numpy().copy()
    out[f"{tag}_z"], out[f"{tag}_y"], out[f"{tag}_alpha"], out[f"{tag}_t"], out[f"{tag}_w"] = z, y, alpha, t, w
np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
print("wrote", len(out), "arrays")
