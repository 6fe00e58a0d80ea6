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
