"""Helpers for the -m gpu tests: build the CUDA model of a golden case and run it."""
import numpy as np
import torch

import fusion_b200 as fb
from fusion_b200 import _lib
from tests.golden import cases as C


def build_model(case, dtype="fp32", flags=0, device="cuda"):
    kw = case["cfg"]
    cfg = C.make_cfg(kw)
    text_model = kw.get("text_model", "one-hot-encoder")
    model = fb.MultimodalModel(cfg.C, cfg.H, device, f"identity:{cfg.F}", text_model, common_dim=cfg.D,
                               text_encoder_dim_output=cfg.T, vocab_size=cfg.V if cfg.V else 91,
                               attention_mecanism=cfg.mechanism, compute_dtype=dtype, engine_flags=flags)
    params = C.gen_params(cfg, case["seed"], np.float32)
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    assert all(k.startswith("text_encoder.") for k in missing.missing_keys) and not missing.unexpected_keys
    return cfg, model.to(device)


def case_inputs(cfg, case, device="cuda"):
    x, tin, labels, cw, masks = C.gen_inputs(cfg, case["B"], case["seed"], case["train"], np.float32)
    tm = None
    if masks is not None:
        tm = {k: torch.from_numpy(v).to(device) for k, v in masks.items()}
    return (torch.from_numpy(x).to(device), torch.from_numpy(tin).to(device), torch.from_numpy(labels).to(device),
            torch.from_numpy(cw).to(device), tm)


def tie_free_inputs(cfg, params64, B, seed, train=True, min_margin=2e-5, rounds=8, forward=None):
    """Seeded float64 inputs of a BIG batch with no ReLU pre-activation within `min_margin` of zero.

    A pre-activation at rounding distance of zero makes that sample's gradient discontinuous (fp32 and float64
    decide the ReLU differently and one sample's whole contribution flips) - a numerical tie, not a parity error.
    With 4096 rows x ~1500 ReLU units some row always sits on one, so the rows the float64 oracle reports as tied
    are redrawn (inputs only: labels, class weights and dropout masks stay) until none is left.
    `forward`: oracle variant that defines the margins (the bf16-rounding oracle for the bf16 tests)."""
    from oracle import head_oracle as ho
    fwd = forward or ho.head_forward_backward
    x, tin, labels, cw, masks = C.gen_inputs(cfg, B, seed, train, np.float64)
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    R = min(64, max(1, 4096 // B))                 # candidates per row and round (pool row i stands in for row i % B)
    pool_masks = None if masks is None else {k: np.tile(v, (R, 1)) for k, v in masks.items()}
    for _ in range(rounds):
        rows = fwd(cfg, params64, x, tin, None, None, masks)["relu_margin_rows"]
        bad = np.nonzero(rows < min_margin)[0] if rows is not None else np.zeros(0, np.int64)
        if bad.size == 0:
            return x, tin, labels, cw, masks
        px = rng.standard_normal((B * R, x.shape[1])); pt = rng.standard_normal((B * R, tin.shape[1]))
        prow = fwd(cfg, params64, px, pt, None, None, pool_masks)["relu_margin_rows"].reshape(R, B)
        for b in bad:
            ok = np.nonzero(prow[:, b] >= min_margin)[0]
            if ok.size:
                x[b] = px[ok[0] * B + b]; tin[b] = pt[ok[0] * B + b]
    raise AssertionError(f"could not draw a tie-free batch in {rounds} rounds ({bad.size} rows left)")


def run_autograd_arrays(model, x, tin, labels, cw, masks, train=True, device="cuda"):
    """Same call shape as run_autograd on explicit numpy inputs (any float dtype; cast to fp32 for the device)."""
    xt = torch.from_numpy(np.asarray(x, np.float32)).to(device).requires_grad_(True)
    tt = torch.from_numpy(np.asarray(tin, np.float32)).to(device)
    y = torch.from_numpy(labels).to(device)
    cwt = torch.from_numpy(np.asarray(cw, np.float32)).to(device)
    tm = None if masks is None else {k: torch.from_numpy(v).to(device) for k, v in masks.items()}
    model.train(train)
    model.inject_dropout_masks(tm)
    model.zero_grad(set_to_none=True)
    logits = model(xt, tt)
    loss = fb.FusedCrossEntropyLoss(weight=cwt)(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: (None if p.grad is None else p.grad.detach().cpu().numpy()) for k, p in model.named_parameters()
             if not k.startswith(("text_encoder.", "image_encoder."))}
    return logits.detach().cpu().numpy(), float(loss), grads, xt.grad.detach().cpu().numpy()


def run_autograd(model, cfg, case, device="cuda"):
    """model(x, meta) -> FusedCrossEntropyLoss -> backward, exactly the call shape of train_pad_20.py:110-112."""
    x, tin, y, cw, masks = case_inputs(cfg, case, device)
    model.train(case["train"])
    model.inject_dropout_masks(masks)
    model.zero_grad(set_to_none=True)
    x.requires_grad_(True)
    logits = model(x, tin)
    loss = fb.FusedCrossEntropyLoss(weight=cw)(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: (None if p.grad is None else p.grad.detach().cpu().numpy()) for k, p in model.named_parameters()
             if not k.startswith(("text_encoder.", "image_encoder."))}
    return logits.detach().cpu().numpy(), float(loss), grads, x.grad.detach().cpu().numpy()


def run_fused(model, cfg, case, device="cuda"):
    """model.forward_loss: forward + CE + backward in one library call."""
    x, tin, y, cw, masks = case_inputs(cfg, case, device)
    model.train(case["train"])
    model.inject_dropout_masks(masks)
    model.zero_grad(set_to_none=True)
    loss, logits = model.forward_loss(x, tin, y, cw)
    torch.cuda.synchronize()
    grads = {k: (None if p.grad is None else p.grad.detach().cpu().numpy()) for k, p in model.named_parameters()
             if not k.startswith(("text_encoder.", "image_encoder."))}
    return logits.cpu().numpy(), float(loss), grads
