"""Fused Adam for the head (and anything else that lives in fp32 on the GPU): one multi-tensor kernel launch
per 48 tensors instead of torch's per-op foreach chain.  Drop-in for the reference's
``optim.Adam(model.parameters(), lr=5e-5, weight_decay=1e-4)`` (train_pad_20.py:54): same constructor arguments,
``param_groups`` (ReduceLROnPlateau keeps working, :55-61), ``state_dict`` keys (``step``, ``exp_avg``,
``exp_avg_sq``) and the same rule that parameters whose ``.grad`` is None are skipped."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.Fb200Error(-2, "FusedAdam handles contiguous float32 CUDA parameters only")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                by_step.setdefault(st["step"], []).append((p, p.grad if p.grad.is_contiguous() else p.grad.contiguous(), st))
            for step, items in by_step.items():
                n = len(items)
                arr = lambda vals: (C.c_void_p * n)(*vals)
                numel = (C.c_int64 * n)(*[p.numel() for p, _, _ in items])
                with torch.cuda.device(items[0][0].device):
                    _lib.check(L.fb200_adam_step(
                        n, arr([p.data_ptr() for p, _, _ in items]), arr([g.data_ptr() for _, g, _ in items]),
                        arr([s["exp_avg"].data_ptr() for _, _, s in items]), arr([s["exp_avg_sq"].data_ptr() for _, _, s in items]),
                        numel, group["lr"], group["betas"][0], group["betas"][1], group["eps"], group["weight_decay"], step,
                        float(grad_scale), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb200_adam_step")
        return loss
