"""CPU oracle of the reference's TabTransformer (TEST INFRASTRUCTURE - never imported by the product path).

Restates, in numpy float64, ``models/tab_transformer.py:6-60`` of the reference:

* ``:10-12``  one ``nn.Embedding(cardinality, embed_dim)`` per categorical column; ``:42-43`` lookups stacked to [B, T, D];
* ``:19-27``  ``nn.TransformerEncoder`` of ``nn.TransformerEncoderLayer(d_model, nhead, dim_feedforward, relu, dropout,
  batch_first=True)`` - torch's post-norm layer: ``x = norm1(x + dropout1(self_attn(x)))``,
  ``x = norm2(x + dropout2(linear2(dropout(relu(linear1(x))))))``, the attention itself dropping normalised
  probabilities with the same ``p`` (``torch/nn/modules/transformer.py`` ``_sa_block`` / ``_ff_block``);
* ``:47``     flatten to [B, T*D]; ``:50-51`` ``numeric_projection``; ``:57`` concatenation; ``:33-38`` the ``fc`` MLP.

Parity is PINNED: ``tests/test_oracle_tabt.py`` checks this file against the unmodified reference class (imported from
/root/reference when present) and against committed golden vectors produced from it (``tests/golden/tabt.npz``,
``tests/golden/make_golden_tabt.py``), in eval mode and in train mode with injected dropout masks.
"""
import numpy as np

from oracle import head_oracle as ho

LAYER_KEYS = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
              "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
              "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")


def layer_shapes(D, F):
    return {"self_attn.in_proj_weight": (3 * D, D), "self_attn.in_proj_bias": (3 * D,), "self_attn.out_proj.weight": (D, D),
            "self_attn.out_proj.bias": (D,), "linear1.weight": (F, D), "linear1.bias": (F,), "linear2.weight": (D, F),
            "linear2.bias": (D,), "norm1.weight": (D,), "norm1.bias": (D,), "norm2.weight": (D,), "norm2.bias": (D,)}


def param_shapes(cards, num_continuous, D, F, L, output_dim):
    """name -> shape in the reference module's state_dict order (tab_transformer.py:10-38)."""
    s = {}
    for i, c in enumerate(cards):
        s[f"embeddings.{i}.weight"] = (c, D)
    for l in range(L):
        for k, shp in layer_shapes(D, F).items():
            s[f"transformer_encoder.layers.{l}.{k}"] = shp
    if num_continuous > 0:
        s["numeric_projection.weight"] = (D, num_continuous)
        s["numeric_projection.bias"] = (D,)
    width = len(cards) * D + (D if num_continuous > 0 else 0)
    s["fc.0.weight"] = (F, width); s["fc.0.bias"] = (F,)
    s["fc.3.weight"] = (output_dim, F); s["fc.3.bias"] = (output_dim,)
    return s


def gen_params(shapes, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in shapes.items():
        if "norm" in k and k.endswith("weight"):
            out[k] = (1.0 + 0.2 * rng.standard_normal(shp)).astype(dtype)
        elif len(shp) == 1:
            out[k] = (0.1 * rng.standard_normal(shp)).astype(dtype)
        elif k.startswith("embeddings"):
            out[k] = rng.standard_normal(shp).astype(dtype)
        else:
            out[k] = (rng.standard_normal(shp) / np.sqrt(shp[1])).astype(dtype)
    return out


def embed(t, tables, x_cat):
    """torch.stack([emb_i(x_cat[:, i])], dim=1) (tab_transformer.py:42-43)."""
    B, T = x_cat.shape
    y = ho.Var(np.stack([tables[i].v[x_cat[:, i]] for i in range(T)], axis=1))

    def bwd():
        if y.grad is None:
            return
        for i in range(T):
            g = np.zeros_like(tables[i].v)
            np.add.at(g, x_cat[:, i], y.grad[:, i])
            tables[i].acc(g)

    t.push(bwd)
    return y


def self_attention(t, x, in_w, in_b, out_w, out_b, H, p, mask):
    """nn.MultiheadAttention(batch_first=True, dropout=p) self-attention on x [B, T, D]; ``mask`` [B, H, T, T] keep-mask on the
    normalised probabilities (None: eval)."""
    B, T, D = x.v.shape
    hd = D // H
    scale = 1.0 / np.sqrt(hd)
    QKV = x.v @ in_w.v.T + in_b.v
    Qh, Kh, Vh = (QKV[..., i * D:(i + 1) * D].reshape(B, T, H, hd).transpose(0, 2, 1, 3) for i in range(3))
    S = (Qh * scale) @ Kh.transpose(0, 1, 3, 2)
    S = S - S.max(axis=-1, keepdims=True)
    P = np.exp(S)
    P = P / P.sum(axis=-1, keepdims=True)
    m = np.ones_like(P) if mask is None else mask.astype(P.dtype) / (1.0 - p)
    Pd = P * m
    O = (Pd @ Vh).transpose(0, 2, 1, 3).reshape(B, T, D)
    y = ho.Var(O @ out_w.v.T + out_b.v)

    def bwd():
        if y.grad is None:
            return
        dy = y.grad
        out_w.acc(dy.reshape(-1, D).T @ O.reshape(-1, D))
        out_b.acc(dy.reshape(-1, D).sum(axis=0))
        dOh = (dy @ out_w.v).reshape(B, T, H, hd).transpose(0, 2, 1, 3)
        dVh = Pd.transpose(0, 1, 3, 2) @ dOh
        dP = (dOh @ Vh.transpose(0, 1, 3, 2)) * m
        dS = P * (dP - (dP * P).sum(axis=-1, keepdims=True))
        dQh = (dS @ Kh) * scale
        dKh = dS.transpose(0, 1, 3, 2) @ (Qh * scale)
        dQKV = np.concatenate([g.transpose(0, 2, 1, 3).reshape(B, T, D) for g in (dQh, dKh, dVh)], axis=-1)
        in_w.acc(dQKV.reshape(-1, 3 * D).T @ x.v.reshape(-1, D))
        in_b.acc(dQKV.reshape(-1, 3 * D).sum(axis=0))
        x.acc(dQKV @ in_w.v)

    t.push(bwd)
    return y


def relu(t, x):
    """ho.relu, also recording the smallest |pre-activation| per SAMPLE in t.sample_margin: a pre-activation within fp32
    rounding of zero makes the gradient discontinuous, and tests redraw such samples instead of comparing coin flips."""
    m = np.abs(x.v).reshape(x.v.shape[0], -1).min(axis=1)
    t.sample_margin = m if getattr(t, "sample_margin", None) is None else np.minimum(t.sample_margin, m)
    return ho.relu(t, x)


def encoder_layer(t, x, P, H, p, masks, l):
    """One post-norm nn.TransformerEncoderLayer; P maps LAYER_KEYS -> Var; masks None (eval) or dict of keep-masks."""
    mk = (lambda k: None) if masks is None else (lambda k: masks[k][l])
    a = self_attention(t, x, P["self_attn.in_proj_weight"], P["self_attn.in_proj_bias"], P["self_attn.out_proj.weight"],
                       P["self_attn.out_proj.bias"], H, p, mk("attn"))
    x1 = ho.layernorm(t, ho.add(t, x, ho.dropout(t, a, p, mk("res1"))), P["norm1.weight"], P["norm1.bias"])
    h = ho.dropout(t, relu(t, ho.linear(t, x1, P["linear1.weight"], P["linear1.bias"])), p, mk("ff"))
    z = ho.linear(t, h, P["linear2.weight"], P["linear2.bias"])
    return ho.layernorm(t, ho.add(t, x1, ho.dropout(t, z, p, mk("res2"))), P["norm2.weight"], P["norm2.bias"])


def flatten(t, x):
    y = ho.Var(x.v.reshape(x.v.shape[0], -1))

    def bwd():
        if y.grad is not None:
            x.acc(y.grad.reshape(x.v.shape))

    t.push(bwd)
    return y


def forward_backward(params, x_cat, x_num, num_heads, p=0.3, masks=None, dout=None, encoder_only=False, d_features=None):
    """TabTransformer.forward (tab_transformer.py:40-60) and its gradients.

    params: name -> array (state_dict names); masks: None (eval) or {"attn" [L,B,H,T,T], "res1" [L,B,T,D], "ff" [L,B,T,F],
    "res2" [L,B,T,D], "fc" [B,F]} keep-masks; returns {"out", "features", "grads": name -> array, "d_num"}."""
    t = ho.Tape()
    t.sample_margin = None
    V = {k: ho.Var(np.asarray(v)) for k, v in params.items()}
    T = x_cat.shape[1]
    L = 1 + max(int(k.split(".")[2]) for k in params if k.startswith("transformer_encoder.layers."))
    x = embed(t, [V[f"embeddings.{i}.weight"] for i in range(T)], x_cat)
    for l in range(L):
        P = {k: V[f"transformer_encoder.layers.{l}.{k}"] for k in LAYER_KEYS}
        x = encoder_layer(t, x, P, num_heads, p, masks, l)
    feats = flatten(t, x)
    xn = None
    if "numeric_projection.weight" in V and not encoder_only:
        xn = ho.Var(np.asarray(x_num, dtype=feats.v.dtype))
        feats = ho.cat(t, feats, ho.linear(t, xn, V["numeric_projection.weight"], V["numeric_projection.bias"]))
    if encoder_only:
        out = feats
    else:
        h = ho.dropout(t, relu(t, ho.linear(t, feats, V["fc.0.weight"], V["fc.0.bias"])), p, None if masks is None else masks["fc"])
        out = ho.linear(t, h, V["fc.3.weight"], V["fc.3.bias"])
    res = {"out": out.v, "features": feats.v, "relu_margin": t.relu_margin, "sample_margin": t.sample_margin}
    g = dout if dout is not None else d_features
    if g is not None:
        out.grad = np.asarray(g, dtype=out.v.dtype).copy()
        t.backward()
        res["grads"] = {k: (v.grad if v.grad is not None else np.zeros_like(v.v)) for k, v in V.items()}
        res["d_num"] = None if xn is None else xn.grad
    return res


def gen_masks(rng, L, B, T, D, F, H, p):
    keep = lambda *s: (rng.random(s) >= p).astype(np.uint8)
    return {"attn": keep(L, B, H, T, T), "res1": keep(L, B, T, D), "ff": keep(L, B, T, F), "res2": keep(L, B, T, D), "fc": keep(B, F)}
