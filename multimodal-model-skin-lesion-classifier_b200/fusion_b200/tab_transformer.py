"""TabTransformer on the fused sm_100a kernels behind the reference's constructor and ``state_dict`` names.

Reference: ``models/tab_transformer.py:6-60`` (built by ``models/loadImageModelClassifier.py:186-198`` as
``TabTransformer([10] * 82, num_continuous=4, output_dim=85)``).  Parameter containers are the reference's own
(``embeddings`` ModuleList of ``nn.Embedding``, ``transformer_encoder`` = ``nn.TransformerEncoder`` of post-norm layers,
``numeric_projection``, ``fc`` Sequential) so checkpoints load either way; ``forward`` never calls those modules:

* embedding lookups + stack + every encoder layer + flatten: ONE launch per pass (``fb200_tabt_forward`` /
  ``fb200_tabt_backward``, csrc/tabt.cu: one sample per CTA, activations in shared memory, backward recomputes);
* numeric projection and the fc MLP: ``fb200_linear_*`` (tcgen05 3xTF32 / FFMA GEMMs); the projection writes straight into the
  right-hand columns of the feature matrix the encoder kernel fills, so there is no concatenation.

There is no CPU path: host tensors raise ``Fb200Error`` (FB200_EUNSUPPORTED).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib

_LAYER_KEYS = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
               "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
               "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(**tensors):
    for name, t in tensors.items():
        if t is not None and not t.is_cuda:
            raise _lib.Fb200Error(-2, f"{name} must be a CUDA tensor (fusion_b200 has no CPU path)")


def _mask_array(masks):
    arr = (C.c_void_p * 4)()
    for i in range(4):
        m = None if masks is None else masks[i]
        arr[i] = m.data_ptr() if m is not None else None
    return arr


class _PackParameters(torch.autograd.Function):
    """flat = cat(p.reshape(-1) for p in params).  torch.cat's own backward launches one copy per parameter (112 of them for the
    reference's TabTransformer - a millisecond of host time, more than the encoder kernels at the reference's batch of 32); this
    one hands every parameter a VIEW of the flat gradient, which autograd installs as .grad without touching the device."""

    @staticmethod
    def forward(ctx, *params):
        ctx.shapes = [p.shape for p in params]
        return torch.cat([p.reshape(-1) for p in params])

    @staticmethod
    def backward(ctx, g):
        outs, off = [], 0
        for shp in ctx.shapes:
            n = shp.numel()
            outs.append(g[off:off + n].view(shp))
            off += n
        return tuple(outs)


class FusedTabEncoderFunction(torch.autograd.Function):
    """features [B, T*D (+ D)] = [flatten(encoder(embed(x_cat))) | numeric_projection(x_num)]."""

    @staticmethod
    def forward(ctx, x_cat, emb_base, flat, x_num, num_w, num_b, cfg):
        L = _lib.lib()
        T, D, H, F, NL, n_rows, train, p, seed, offset, masks = cfg
        _need_cuda(x_categorical=x_cat, parameters=flat, x_numerical=x_num if num_w is not None else None)
        if x_cat.dtype != torch.int64 or x_cat.dim() != 2 or x_cat.shape[1] != T:
            raise ValueError(f"x_categorical must be an int64 [B, {T}] tensor")
        x_cat = x_cat.contiguous()
        B = x_cat.shape[0]
        desc = _lib.TabtDesc(B=B, T=T, D=D, H=H, F=F, L=NL, n_emb_rows=n_rows, train=int(train), p=float(p), flags=0)
        sb, wb = C.c_size_t(0), C.c_size_t(0)
        _lib.check(L.fb200_tabt_workspace_bytes(C.byref(desc), C.byref(sb), C.byref(wb)), "fb200_tabt_workspace_bytes")
        dev = flat.device
        width = T * D + (D if num_w is not None else 0)
        feats = torch.empty(B, width, dtype=torch.float32, device=dev)
        saved = torch.empty(sb.value, dtype=torch.uint8, device=dev)
        marr = _mask_array(masks)
        with torch.cuda.device(dev):
            _lib.check(L.fb200_tabt_forward(C.byref(desc), _ptr(x_cat), _ptr(emb_base), _ptr(flat), marr, seed, offset, None,
                                            _ptr(feats), width, _ptr(saved), _stream()), "fb200_tabt_forward")
            if num_w is not None:
                xn = x_num.to(torch.float32).contiguous()
                nc = xn.shape[1]
                _lib.check(L.fb200_linear_forward(B, D, nc, _ptr(xn), nc, _ptr(num_w), _ptr(num_b), 0,
                                                  C.c_void_p(feats.data_ptr() + 4 * T * D), width, _stream()), "fb200_linear_forward")
            else:
                xn = None
        ctx.desc, ctx.cfg, ctx.ws_bytes = desc, cfg, wb.value
        ctx.save_for_backward(x_cat, emb_base, flat, saved, xn, num_w)
        return feats

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dfeats):
        L = _lib.lib()
        x_cat, emb_base, flat, saved, xn, num_w = ctx.saved_tensors
        T, D, H, F, NL, n_rows, train, p, seed, offset, masks = ctx.cfg
        desc = ctx.desc
        dfeats = dfeats.to(torch.float32).contiguous()
        width = dfeats.shape[1]
        dflat = torch.empty_like(flat)
        ws = torch.empty(ctx.ws_bytes, dtype=torch.uint8, device=flat.device)
        marr = _mask_array(masks)
        d_xn = d_nw = d_nb = None
        with torch.cuda.device(flat.device):
            _lib.check(L.fb200_tabt_backward(C.byref(desc), _ptr(x_cat), _ptr(emb_base), _ptr(flat), marr, seed, offset, None,
                                             _ptr(saved), _ptr(dfeats), width, _ptr(dflat), _ptr(ws), _stream()), "fb200_tabt_backward")
            if num_w is not None:
                B, nc = xn.shape
                d_nw, d_nb = torch.empty_like(num_w), torch.empty(D, dtype=torch.float32, device=flat.device)
                d_xn = torch.empty_like(xn) if ctx.needs_input_grad[3] else None
                _lib.check(L.fb200_linear_backward(B, D, nc, _ptr(xn), nc, _ptr(num_w), C.c_void_p(dfeats.data_ptr() + 4 * T * D), width,
                                                   _ptr(d_xn), nc, _ptr(d_nw), _ptr(d_nb), _stream()), "fb200_linear_backward")
        return None, None, dflat, d_xn, d_nw, d_nb, None


class FusedLinearFunction(torch.autograd.Function):
    """y = x W^T + b (optionally ReLU) on the library's GEMM engines; x [M, K] fp32 contiguous."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        L = _lib.lib()
        _need_cuda(x=x, weight=w, bias=b)
        x = x.to(torch.float32).contiguous()
        w, b = w.contiguous(), b.contiguous()
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.fb200_linear_forward(M, N, K, _ptr(x), K, _ptr(w), _ptr(b), int(relu), _ptr(y), N, _stream()), "fb200_linear_forward")
        ctx.relu = relu
        ctx.save_for_backward(x, w, y if relu else None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        L = _lib.lib()
        x, w, y = ctx.saved_tensors
        dy = dy.to(torch.float32)
        dy = (dy * (y > 0)) if ctx.relu else dy.contiguous()
        M, K = x.shape
        N = w.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, db = torch.empty_like(w), torch.empty(N, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.fb200_linear_backward(M, N, K, _ptr(x), K, _ptr(w), _ptr(dy), N, _ptr(dx), K, _ptr(dw), _ptr(db), _stream()),
                       "fb200_linear_backward")
        return dx, dw, db, None


class TabTransformer(nn.Module):
    """Drop-in for the reference's ``TabTransformer`` (same constructor, same ``state_dict`` keys, same outputs)."""

    def __init__(self, categorical_cardinalities, num_continuous, embed_dim=32, num_heads=4, num_transformer_layers=2,
                 hidden_dim=128, output_dim=1, dropout=0.3):
        super().__init__()
        self.embeddings = nn.ModuleList([nn.Embedding(c, embed_dim) for c in categorical_cardinalities])
        self.num_categorical, self.embed_dim = len(categorical_cardinalities), embed_dim
        layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=num_heads, dim_feedforward=hidden_dim,
                                           activation="relu", dropout=dropout, batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=num_transformer_layers)
        self.numeric_projection = nn.Linear(num_continuous, embed_dim) if num_continuous > 0 else None
        width = self.num_categorical * embed_dim + (embed_dim if num_continuous > 0 else 0)
        self.fc = nn.Sequential(nn.Linear(width, hidden_dim), nn.ReLU(), nn.Dropout(dropout), nn.Linear(hidden_dim, output_dim))
        self.num_heads, self.hidden_dim, self.num_layers, self.p = num_heads, hidden_dim, num_transformer_layers, float(dropout)
        base, acc = [], 0
        for c in categorical_cardinalities:
            base.append(acc)
            acc += int(c)
        self.n_emb_rows = acc
        self.register_buffer("_emb_base", torch.tensor(base, dtype=torch.int32), persistent=False)
        self._rng_calls = 0
        self._seed = (int(torch.initial_seed()) ^ 0x7AB7) & 0x7FFFFFFFFFFFFFFF      # torch.manual_seed controls the dropout stream
        self._flat_cache = None
        self._test_masks = None          # tests inject {"enc": (attn, res1, ff, res2) uint8 tensors, "fc": uint8 [B, hidden]}

    def _ordered_params(self):
        ps = []
        for layer in self.transformer_encoder.layers:
            named = dict(layer.named_parameters())
            ps += [named[k] for k in _LAYER_KEYS]
        return ps + [e.weight for e in self.embeddings]

    def _flat_params(self):
        """The kernels' parameter layout (include/fb200.h): layer blocks, then the stacked embedding tables, packed by a function
        whose backward hands every parameter a view of the flat gradient.  The packed copy is reused while no parameter changed
        (eval loops, frozen encoders): the key is every parameter's (storage, version)."""
        ps = self._ordered_params()
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in ps)):
            key = tuple((p.data_ptr(), p._version) for p in ps)
            if self._flat_cache is not None and self._flat_cache[0] == key:
                return self._flat_cache[1]
            flat = torch.cat([p.detach().reshape(-1) for p in ps])
            self._flat_cache = (key, flat)
            return flat
        return _PackParameters.apply(*ps)

    def encode(self, x_categorical, x_numerical=None):
        """[flatten(transformer_encoder(stack(embeddings))) | numeric_projection(x_numerical)]  (tab_transformer.py:42-57)."""
        train = self.training and self.p > 0.0
        offset = self._rng_calls
        if train:
            self._rng_calls += 1
        masks = None if self._test_masks is None else self._test_masks["enc"]
        cfg = (self.num_categorical, self.embed_dim, self.num_heads, self.hidden_dim, self.num_layers, self.n_emb_rows,
               train, self.p, self._seed, offset, masks)
        nw = nb = None
        if self.numeric_projection is not None:
            nw, nb = self.numeric_projection.weight, self.numeric_projection.bias
        return FusedTabEncoderFunction.apply(x_categorical, self._emb_base, self._flat_params(), x_numerical, nw, nb, cfg)

    def forward(self, x_categorical, x_numerical):
        feats = self.encode(x_categorical, x_numerical)
        h = FusedLinearFunction.apply(feats, self.fc[0].weight, self.fc[0].bias, True)
        if self.training and self.p > 0.0:
            if self._test_masks is not None:
                h = h * (self._test_masks["fc"].to(h.dtype) / (1.0 - self.p))
            else:
                h = torch.nn.functional.dropout(h, self.p, True)            # [B, hidden_dim]: the one stock element-wise op left
        return FusedLinearFunction.apply(h, self.fc[3].weight, self.fc[3].bias, False)
