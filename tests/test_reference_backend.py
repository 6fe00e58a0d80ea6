"""backend="reference" (SURVEY.md 8b caveat): the stock-torch composition of the head for callers that differentiate twice
(Grad-CAM++: src/services/XAI/models/cam.py:38-43).  CPU: every fusion string against the float64 oracle, and a double
backward.  GPU: same logits / gradients as the fused CUDA path."""
import numpy as np
import pytest
import torch

import fusion_b200 as fb
from oracle import head_oracle as ho
from tests import parity
from tests.golden import cases as C


def _model(mech, dims, device="cpu", backend="reference", dtype=torch.float64, seed=3):
    cfg = C.make_cfg(dict(dims, mechanism=mech))
    m = fb.MultimodalModel(cfg.C, cfg.H, device, f"identity:{cfg.F}", "one-hot-encoder", common_dim=cfg.D, text_encoder_dim_output=cfg.T,
                           vocab_size=cfg.V, attention_mecanism=mech, backend=backend)
    params = C.gen_params(cfg, seed, np.float64)
    m = m.to(device=device, dtype=dtype)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    return cfg, m, params


@pytest.mark.parametrize("mech", ho.MECHANISMS)
def test_reference_backend_matches_oracle_on_cpu(mech):
    cfg, m, params = _model(mech, C.SMALL_DIMS)
    x, tin, labels, cw, _ = C.gen_inputs(cfg, 7, 5, False, np.float64)
    m.eval()
    xt = torch.from_numpy(x).requires_grad_(True)
    logits = m(xt, torch.from_numpy(tin))
    loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(labels), weight=torch.from_numpy(cw))
    loss.backward()
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, None, need_input_grad=True)
    assert parity.rel_err(logits.detach().numpy(), o["logits"]) < 1e-12
    for k, g in o["grads"].items():
        p = dict(m.named_parameters())[k]
        if g is None:
            assert p.grad is None, k
        else:
            assert parity.rel_err(p.grad.numpy(), g) < 1e-10, k
    assert parity.rel_err(xt.grad.numpy(), o["d_img_feat"]) < 1e-10


def test_reference_backend_is_twice_differentiable():
    cfg, m, _ = _model("crossattention", C.SMALL_DIMS)
    x, tin, labels, cw, _ = C.gen_inputs(cfg, 4, 9, False, np.float64)
    xt = torch.from_numpy(x).requires_grad_(True)
    score = m.eval()(xt, torch.from_numpy(tin))[:, 1].sum()
    (g1,) = torch.autograd.grad(score, xt, create_graph=True)            # cam.py:38-43 pattern
    (g2,) = torch.autograd.grad(g1.pow(2).sum(), xt)
    assert torch.isfinite(g2).all() and g2.abs().max() > 0
    with pytest.raises(ValueError):
        fb.MultimodalModel(2, 8, "cpu", "identity:8", "one-hot-encoder", backend="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("mech", ["crossattention", "metablock", ho.RG_ATT, "gfcam"])
def test_reference_backend_equals_fused_path_on_gpu(mech):
    dims = dict(F=512, V=85, C=6)
    cfg, fused, _ = _model(mech, dims, "cuda", "b200", torch.float32)
    x, tin, labels, cw, _ = C.gen_inputs(cfg, 48, 5, False, np.float32)
    xt, tt = torch.from_numpy(x).cuda(), torch.from_numpy(tin).cuda()
    y, w = torch.from_numpy(labels).cuda(), torch.from_numpy(cw).cuda()
    fused.eval()
    out = {}
    for backend in ("b200", "reference"):
        fused.backend = backend                      # same parameters, same state dict
        fused.zero_grad(set_to_none=True)
        logits = fused(xt, tt)
        torch.nn.functional.cross_entropy(logits, y, weight=w).backward()
        out[backend] = (logits.detach().cpu().numpy(), {k: None if p.grad is None else p.grad.cpu().numpy() for k, p in fused.named_parameters()})
    assert parity.rel_err(out["b200"][0], out["reference"][0]) < 1e-5
    for k, g in out["reference"][1].items():
        if k.startswith(("image_encoder.", "text_encoder.")):
            continue
        if g is None:
            assert out["b200"][1][k] is None, k
        else:
            assert parity.rel_err(out["b200"][1][k], g) < 2e-5, k
