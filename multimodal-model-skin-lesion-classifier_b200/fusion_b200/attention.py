"""Multi-head attention over token sequences through the C ABI (fb200_mha_forward / fb200_mha_backward).

Drop-in for ``torch.nn.MultiheadAttention(embed_dim, num_heads)`` (``batch_first=False``, ``dropout=0``) as the
reference builds it (models/multimodalIntraInterModal.py:78-100) and as its sequence variants call it with real token
sequences - image tokens attending to metadata tokens, models/multimodalGated.py:118-206: same constructor
arguments, same parameter names (``in_proj_weight``, ``in_proj_bias``, ``out_proj.weight``, ``out_proj.bias`` - state
dicts interchange), same initialisation, ``forward(query, key, value) -> (output, None)``.  The head-averaged
attention weights are not produced: every call site of the reference discards them.

All arithmetic runs in the CUDA kernels of csrc/attention.cuh (fused softmax(QK^T)V with the probabilities held
on chip, flash-style recomputation in backward) and the library's GEMM engines; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from .head import _ptr, _stream


def _f32c(t, name):
    if not t.is_cuda:
        raise _lib.Fb200Error(-2, f"{name} must be a CUDA tensor (fusion_b200 has no CPU path)")
    return t.detach().to(torch.float32).contiguous()


class FusedMHAFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, key, value, in_w, in_b, out_w, out_b, num_heads, pool_mean=False):
        L = _lib.lib()
        q, k, v = _f32c(query, "query"), _f32c(key, "key"), _f32c(value, "value")
        if q.dim() != 3 or k.dim() != 3 or v.shape != k.shape or q.shape[1:] != k.shape[1:]:
            raise ValueError(f"expected query [Sq,B,D], key/value [Skv,B,D]; got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
        Sq, B, D = q.shape
        if D % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")       # torch's own check, same exception type
        desc = _lib.MhaDesc(Sq=Sq, Skv=k.shape[0], B=B, D=D, H=num_heads, flags=1 if pool_mean else 0)     # FB200_MHA_POOL_MEAN
        nbytes = C.c_size_t(0)
        _lib.check(L.fb200_mha_workspace_bytes(C.byref(desc), C.byref(nbytes)), "fb200_mha_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=q.device)
        out = torch.empty(B, D, dtype=torch.float32, device=q.device) if pool_mean else torch.empty_like(q)
        w = [_f32c(t, n) for t, n in ((in_w, "in_proj_weight"), (in_b, "in_proj_bias"), (out_w, "out_proj.weight"), (out_b, "out_proj.bias"))]
        with torch.cuda.device(q.device):
            _lib.check(L.fb200_mha_forward(C.byref(desc), _ptr(q), _ptr(k), _ptr(v), _ptr(w[0]), _ptr(w[1]), _ptr(w[2]), _ptr(w[3]),
                                           _ptr(out), _ptr(ws), _stream()), "fb200_mha_forward")
        ctx.desc, ctx.ws, ctx.qkv, ctx.w = desc, ws, (q, k, v), w
        ctx.need = tuple(bool(t.requires_grad) for t in (query, key, value))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        L = _lib.lib()
        q, k, v = ctx.qkv
        w = ctx.w
        do = _f32c(dout, "dout")
        dq = torch.empty_like(q) if ctx.need[0] else None
        dk = torch.empty_like(k) if ctx.need[1] else None
        dv = torch.empty_like(v) if ctx.need[2] else None
        d_in_w, d_in_b, d_out_w, d_out_b = (torch.empty_like(t) for t in w)
        with torch.cuda.device(q.device):
            _lib.check(L.fb200_mha_backward(C.byref(ctx.desc), _ptr(q), _ptr(k), _ptr(v), _ptr(w[0]), _ptr(w[2]), _ptr(do),
                                            _ptr(dq), _ptr(dk), _ptr(dv), _ptr(d_in_w), _ptr(d_in_b), _ptr(d_out_w), _ptr(d_out_b),
                                            _ptr(ctx.ws), _stream()), "fb200_mha_backward")
        ctx.ws = None
        return dq, dk, dv, d_in_w, d_in_b, d_out_w, d_out_b, None, None


class MultiheadAttention(nn.Module):
    """``nn.MultiheadAttention(embed_dim, num_heads)`` with the fused CUDA path (sequence-first tensors)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=True, batch_first=False, pool=None):
        """``pool="mean"``: forward returns the mean of the attention output over the query tokens, [B, D] - what the
        reference's sequence models compute right after the module (multimodalGated.py:200-205).  The pooling is folded in
        front of the output projection (it commutes with it), so that projection runs on B rows instead of S_q * B."""
        super().__init__()
        if pool not in (None, "mean"):
            raise ValueError("pool must be None or 'mean'")
        self.pool = pool
        if dropout != 0.0 or not bias or batch_first:
            raise ValueError("the fused attention implements the reference's configuration: dropout=0, bias=True, batch_first=False")
        if embed_dim % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.batch_first = False
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self._reset_parameters()

    def _reset_parameters(self):                   # torch/nn/modules/activation.py: MultiheadAttention._reset_parameters
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.0)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=False, attn_mask=None, **_):
        if key_padding_mask is not None or attn_mask is not None:
            raise ValueError("masks are not supported by the fused attention (the reference never passes one)")
        out = FusedMHAFunction.apply(query, key, value, self.in_proj_weight, self.in_proj_bias,
                                     self.out_proj.weight, self.out_proj.bias, self.num_heads, self.pool == "mean")
        return out, None
