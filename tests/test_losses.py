"""Auxiliary losses: oracle vs golden vectors from the reference classes (CPU), CUDA kernels vs both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo
from tests import parity
from tests.golden import cases as C

G = np.load(os.path.join(C.GOLDEN_DIR, "losses.npz"))
TAGS = ["a", "b", "c"]


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("use_w", [True, False])
def test_oracle_matches_reference_golden(tag, use_w):
    n = "w" if use_w else "n"
    z, y, t = G[f"{tag}_z"], G[f"{tag}_y"], G[f"{tag}_t"]
    loss, dz = lo.focal_loss(z, y, G[f"{tag}_alpha"] if use_w else None, 2.0)
    assert abs(loss - G[f"focal_{tag}{n}_loss"]) < 1e-12 and parity.rel_err(dz, G[f"focal_{tag}{n}_dz"]) < 1e-12
    loss, dz = lo.soft_target_ce(z, t, G[f"{tag}_w"] if use_w else None)
    assert abs(loss - G[f"soft_{tag}{n}_loss"]) < 1e-12 and parity.rel_err(dz, G[f"soft_{tag}{n}_dz"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("use_w", [True, False])
def test_cuda_losses_match_reference_golden(tag, use_w):
    import fusion_b200 as fb
    n = "w" if use_w else "n"
    dev = lambda a, dt=torch.float32: torch.tensor(a, dtype=dt, device="cuda")
    z = dev(G[f"{tag}_z"]).requires_grad_(True)
    loss = fb.FusedFocalLoss(alpha=dev(G[f"{tag}_alpha"]) if use_w else None, gamma=2)(z, dev(G[f"{tag}_y"], torch.int64))
    loss.backward()
    assert abs(float(loss) - G[f"focal_{tag}{n}_loss"]) < 1e-5 * abs(G[f"focal_{tag}{n}_loss"])
    assert parity.rel_err(z.grad.cpu().numpy(), G[f"focal_{tag}{n}_dz"]) < 1e-5
    z = dev(G[f"{tag}_z"]).requires_grad_(True)
    loss = fb.FusedSoftTargetCrossEntropy(weight=dev(G[f"{tag}_w"]) if use_w else None)(z, dev(G[f"{tag}_t"]))
    loss.backward()
    assert abs(float(loss) - G[f"soft_{tag}{n}_loss"]) < 1e-5 * abs(G[f"soft_{tag}{n}_loss"])
    assert parity.rel_err(z.grad.cpu().numpy(), G[f"soft_{tag}{n}_dz"]) < 1e-5


@pytest.mark.gpu
def test_softmax_argmax_eval_tail():
    import fusion_b200 as fb
    rng = np.random.default_rng(1)
    z = rng.standard_normal((1001, 6)).astype(np.float32)
    z[5, 2] = z[5, 4] = 9.0                                   # tie: the first maximum wins, as torch.argmax
    probs, pred = fb.softmax_argmax(torch.from_numpy(z).cuda())
    ref = torch.softmax(torch.from_numpy(z).double(), dim=1)
    assert parity.rel_err(probs.cpu().numpy(), ref.numpy()) < 1e-6
    assert torch.equal(pred.cpu(), torch.argmax(torch.from_numpy(z), dim=1))
