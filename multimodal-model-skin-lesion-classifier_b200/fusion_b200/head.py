"""Host-side driver of the fused fusion head: torch tensors in, raw pointers to libfb200.

PyTorch is used for device memory, streams and autograd bookkeeping only; all arithmetic of
the head runs in the CUDA kernels behind the C ABI (include/fb200.h).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_desc(mechanism, B, F, V, T, D, H, Cn, n=2, text_mode=0, dtype="fp32", train=False, flags=0):
    if isinstance(mechanism, str):
        mid = _lib.mechanism_id(mechanism)
        if mid < 0:
            # same message as multimodalIntraInterModal.py:413-416
            raise ValueError(f"Attention mechanism '{mechanism}' not implemented.")
        mechanism = mid
    dt = {"fp32": _lib.F32, "float32": _lib.F32, "bf16": _lib.BF16, "bfloat16": _lib.BF16}[dtype] if isinstance(dtype, str) else dtype
    return _lib.Desc(mechanism=mechanism, B=B, F=F, V=V or 0, T=T, D=D, H=H, C=Cn, n=n, text_mode=text_mode,
                     dtype=dt, train=1 if train else 0, flags=flags, reserved=0)


def denom_arg(denom, device):
    """The weighted-CE denominator reaches the kernels as a raw pointer to ONE fp32 device scalar: cast and move it here
    (float64 class weights from numpy / sklearn, or a host tensor, would otherwise be read as garbage).  A tensor that
    already is fp32 on `device` is returned as is, so static buffers of captured graphs keep their address."""
    if denom is None:
        return None
    if not isinstance(denom, torch.Tensor):
        denom = torch.tensor([float(denom)], dtype=torch.float32)
    if denom.numel() != 1:
        raise ValueError(f"denom must hold one element, got shape {tuple(denom.shape)}")
    return denom.to(device=device, dtype=torch.float32).reshape(1).contiguous()


def _check_input(t, name, cols):
    if not t.is_cuda:
        raise _lib.Fb200Error(-2, f"{name} must be a CUDA tensor (fusion_b200 has no CPU path)")
    if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != cols:
        raise ValueError(f"{name}: expected float32 [B,{cols}], got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


_table_cache = {}


def param_table(tensors):
    """ParamTable for `tensors`, cached on their device pointers: the eager drop-in route builds one per pass, and
    filling a 78-entry ctypes array costs more host time than the whole persistent-kernel launch it precedes."""
    key = tuple(0 if t is None else t.data_ptr() for t in tensors)
    hit = _table_cache.get(key)
    if hit is None:
        if len(_table_cache) > 64:
            _table_cache.clear()
        hit = _table_cache[key] = ParamTable(tensors)
    else:
        hit.tensors = list(tensors)          # same storage, possibly new tensor objects: keep these alive
    return hit


class ParamTable:
    """Array of the 78 parameter pointers in slot order (NULL for absent slots)."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        arr = (C.c_void_p * len(self.tensors))()
        for i, t in enumerate(self.tensors):
            if t is not None:
                if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.Fb200Error(-2, "parameters must be contiguous float32 CUDA tensors")
                arr[i] = t.data_ptr()
        self.arr = arr


def _mask_table(masks):
    """masks: None or {site index or name: uint8 CUDA tensor} -> (ctypes array | NULL, keepalive)."""
    if not masks:
        return None, None
    arr = (C.c_void_p * _lib.NUM_DROPOUT_SITES)()
    keep = []
    for k, t in masks.items():
        idx = _lib.DROP_SITES.index(k) if isinstance(k, str) else int(k)
        t = t.to(torch.uint8).contiguous()
        if not t.is_cuda:
            raise _lib.Fb200Error(-2, "dropout masks must be CUDA tensors")
        arr[idx] = t.data_ptr()
        keep.append(t)
    return arr, keep


class FusedHeadFunction(torch.autograd.Function):
    """logits = head(img_feat, text_in; params) with a hand-written backward.

    Once differentiable (Grad-CAM++ style double backward is not supported - SURVEY 8b)."""

    @staticmethod
    def forward(ctx, img_feat, text_in, cfg, masks, seed, offset, slots, all_params, *params):
        """`params`: the LIVE parameters only (the ones the fusion string differentiates; `slots` = their slot numbers) -
        autograd's per-argument bookkeeping is host time on the critical path of a 100-microsecond step; `all_params`:
        every slot's tensor (or None), for the pointer table."""
        L = _lib.lib()
        B = img_feat.shape[0]
        need_dimg = bool(img_feat.requires_grad)
        need_dtxt = bool(text_in.requires_grad)
        flags = cfg["flags"] | (_lib.FLAG_NEED_DIMG if need_dimg else 0) | (_lib.FLAG_NEED_DTEXT if need_dtxt else 0)
        desc = make_desc(cfg["mechanism"], B, cfg["F"], cfg["V"], cfg["T"], cfg["D"], cfg["H"], cfg["C"], cfg["n"],
                         cfg["text_mode"], cfg["dtype"], cfg["train"], flags)
        x = _check_input(img_feat.detach(), "img_feat", cfg["F"])
        t = _check_input(text_in.detach(), "text_metadata", cfg["T"] if cfg["text_mode"] else cfg["V"])
        table = param_table(all_params)
        ws = torch.empty(_lib.workspace_bytes(desc), dtype=torch.uint8, device=x.device)
        logits = torch.empty(B, cfg["C"], dtype=torch.float32, device=x.device)
        marr, mkeep = _mask_table(masks)
        with torch.cuda.device(x.device):
            _lib.check(L.fb200_head_forward(C.byref(desc), table.arr, _ptr(x), _ptr(t), marr, seed, offset, None,
                                            _ptr(logits), _ptr(ws), _stream()), "fb200_head_forward")
        ctx.desc, ctx.table, ctx.ws, ctx.x, ctx.t = desc, table, ws, x, t
        ctx.marr, ctx.mkeep, ctx.seed, ctx.offset = marr, mkeep, seed, offset
        ctx.need = (need_dimg, need_dtxt)
        ctx.slots = slots
        return logits

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dlogits):
        L = _lib.lib()
        desc = ctx.desc
        dl = dlogits.contiguous().float()
        total, offs = _lib.grad_layout(desc)
        flat = torch.empty(max(total, 1), dtype=torch.float32, device=dl.device)
        d_img = torch.empty_like(ctx.x) if ctx.need[0] else None
        d_txt = torch.empty_like(ctx.t) if ctx.need[1] else None
        with torch.cuda.device(dl.device):
            _lib.check(L.fb200_head_backward(C.byref(desc), ctx.table.arr, _ptr(ctx.x), _ptr(ctx.t), ctx.marr, ctx.seed, ctx.offset, None,
                                             _ptr(dl), _ptr(flat), _ptr(d_img), _ptr(d_txt), _ptr(ctx.ws), _stream()),
                       "fb200_head_backward")
        grads = []
        for s in ctx.slots:                 # parameters outside `slots` never entered the graph: they keep grad=None like in
            p = ctx.table.tensors[s]        # the reference's autograd
            grads.append(flat[offs[s]: offs[s] + p.numel()].view(p.shape))
        ctx.ws = None
        return (d_img, d_txt, None, None, None, None, None, None, *grads)


def cross_entropy(logits, labels, class_w=None, denom=None, want_grad=True):
    """Fused weighted CE through the C ABI.  Returns (loss_out[3] = loss, numerator, weight sum; dlogits)."""
    L = _lib.lib()
    if not logits.is_cuda:
        raise _lib.Fb200Error(-2, "logits must be a CUDA tensor")
    z = logits.detach().contiguous().float()
    y = labels.to(device=z.device, dtype=torch.int64).contiguous()
    w = None if class_w is None else class_w.to(device=z.device, dtype=torch.float32).contiguous()
    out = torch.empty(3, dtype=torch.float32, device=z.device)
    dl = torch.empty_like(z) if want_grad else None
    denom = denom_arg(denom, z.device)
    with torch.cuda.device(z.device):
        _lib.check(L.fb200_cross_entropy(_ptr(z), _ptr(y), _ptr(w), _ptr(denom), z.shape[0], z.shape[1], _ptr(out), _ptr(dl), _stream()),
                   "fb200_cross_entropy")
    return out, dl


class _FusedCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, class_w, denom):
        out, dl = cross_entropy(logits, labels, class_w, denom, want_grad=True)
        ctx.save_for_backward(dl)
        return out[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dloss):
        (dl,) = ctx.saved_tensors
        return dl * dloss, None, None, None


class FusedCrossEntropyLoss(torch.nn.Module):
    """Drop-in for ``nn.CrossEntropyLoss(weight=class_weights)`` (train_pad_20.py:52):
    same ``criterion(outputs, label)`` call, weighted-mean reduction, one fused kernel pair.
    ``denom`` (device scalar) replaces this batch's weight sum by the global one under DP."""

    def __init__(self, weight=None):
        super().__init__()
        self.register_buffer("weight", None if weight is None else weight.detach().clone().float())

    def forward(self, logits, labels, denom=None):
        return _FusedCEFunction.apply(logits, labels, self.weight, denom)


class _AuxLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, weight, kind, gamma):
        L = _lib.lib()
        if not logits.is_cuda:
            raise _lib.Fb200Error(-2, "logits must be a CUDA tensor")
        z = logits.detach().contiguous().float()
        t = targets.to(device=z.device, dtype=torch.int64 if kind == 1 else torch.float32).contiguous()
        w = None if weight is None else weight.to(device=z.device, dtype=torch.float32).contiguous()
        out = torch.empty(1, dtype=torch.float32, device=z.device)
        dl = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _lib.check(L.fb200_aux_loss(kind, _ptr(z), _ptr(t), _ptr(w), float(gamma), z.shape[0], z.shape[1], _ptr(out), _ptr(dl), _stream()), "fb200_aux_loss")
        ctx.save_for_backward(dl)
        return out[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dloss):
        (dl,) = ctx.saved_tensors
        return dl * dloss, None, None, None, None


class FusedFocalLoss(torch.nn.Module):
    """Drop-in for the reference's FocalLoss(alpha, gamma, reduction='mean') (models/focalLoss.py:6-26)."""

    def __init__(self, alpha=None, gamma=2, reduction="mean"):
        super().__init__()
        if reduction != "mean":
            raise ValueError("the fused focal loss implements reduction='mean' (what the reference's loops use)")
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, inputs, targets):
        return _AuxLossFunction.apply(inputs, targets, self.alpha, 1, self.gamma)


class FusedSoftTargetCrossEntropy(torch.nn.Module):
    """Drop-in for SoftTargetCrossEntropy(weight) (models/softtargetsCrossEntropy.py:5-22)."""

    def __init__(self, weight=None):
        super().__init__()
        self.weight = weight

    def forward(self, inputs, targets):
        return _AuxLossFunction.apply(inputs, targets, self.weight, 2, 0.0)


def softmax_argmax(logits):
    """(probs [B,C], preds [B]) in one kernel - the evaluation tail of utils/model_metrics.py:57-58."""
    L = _lib.lib()
    if not logits.is_cuda:
        raise _lib.Fb200Error(-2, "logits must be a CUDA tensor")
    z = logits.detach().contiguous().float()
    probs = torch.empty_like(z)
    pred = torch.empty(z.shape[0], dtype=torch.int64, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(L.fb200_softmax_argmax(_ptr(z), z.shape[0], z.shape[1], _ptr(probs), _ptr(pred), _stream()), "fb200_softmax_argmax")
    return probs, pred
