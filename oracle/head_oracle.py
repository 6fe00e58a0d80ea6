"""CPU oracle for the multimodal fusion-head hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference's fusion head, used only as a
checker by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` leg of
``bench.py``.  Nothing under the product package may import it.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference module itself, imported unmodified on CPU
(``oracle/ref_shim.py``) - see ``tests/golden/make_golden.py`` (fixtures) and
``tests/test_oracle_vs_reference.py`` (live check when /root/reference exists).

What it follows (paths relative to /root/reference/src/scripts/benchmark/models):
  * multimodalIntraInterModal.py:55-160   layer shapes
  * multimodalIntraInterModal.py:162-200  common prefix (projections, 4 MHAs at S=1)
  * multimodalIntraInterModal.py:205-416  the 18 fusion strings
  * gatedResidualBlock.py:4-17            gated residual + LayerNorm
  * metablock.py:4-32                     MetaBlock
  * torch.nn.MultiheadAttention / LayerNorm / CrossEntropyLoss(weight) semantics
    (torch==2.4.1 pinned by the reference's requirements.txt:3).

The arithmetic runs in the dtype of the arrays handed in (float64 for the tight
1e-5 checks, float32 for timing).  Gradients are produced by a ~100-line reverse
tape so that every fusion string is a straight transcription of the forward pass.
"""
from __future__ import annotations

import numpy as np

LN_EPS = 1e-5

MECHANISMS = (
    "no-metadata",
    "no-metadata-without-mlp",
    "concatenation",
    "crossattention",
    "weighted",
    "gfcam",
    "cross-weights-after-crossattention",
    "metablock",
    "rg-att2fusefeatures",
    "rg-att",
    "att-intramodal",
    "att-intramodal+residual",
    "cross-attention-only",
    "residual+cross-attention-metadados",
    "att-intramodal+residual+cross-attention-metadados",
    "att-intramodal+residual+cross-attention-metadados+rg-att2fusefeatures",
    "att-intramodal+residual+cross-attention-metadados+metablock",
    "att-intramodal+residual+cross-attention-metadados+att-intramodal+residual",
)
RG_ATT = "att-intramodal+residual+cross-attention-metadados"
IN_SCOPE = ("concatenation", "metablock", "crossattention", "weighted", "gfcam", RG_ATT)


# --------------------------------------------------------------------------- tape
class Var:
    """A value on the tape.  ``grad`` stays None until something flows into it."""

    __slots__ = ("v", "grad", "name")

    def __init__(self, v, name=None):
        self.v = v
        self.grad = None
        self.name = name

    def acc(self, g):
        self.grad = g.copy() if self.grad is None else self.grad + g


class Tape:
    def __init__(self):
        self.steps = []
        self.relu_margin = np.inf      # smallest |pre-activation| any ReLU saw (ties make gradients discontinuous)
        self.relu_margin_rows = None   # the same per batch row (tests redraw the rows that sit on a tie)

    def push(self, fn):
        self.steps.append(fn)

    def backward(self):
        for fn in reversed(self.steps):
            fn()


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def linear(t, x, W, b):
    """nn.Linear: y = x W^T + b.  W is [out, in]."""
    y = Var(x.v @ W.v.T + b.v)

    def bwd():
        if y.grad is None:
            return
        x.acc(y.grad @ W.v)
        g2 = y.grad.reshape(-1, y.grad.shape[-1])
        W.acc(g2.T @ x.v.reshape(-1, x.v.shape[-1]))
        b.acc(g2.sum(axis=0))

    t.push(bwd)
    return y


def relu(t, x):
    y = Var(np.maximum(x.v, 0))
    if x.v.size:
        rows = np.abs(x.v).min(axis=-1).reshape(-1)
        t.relu_margin = min(t.relu_margin, float(rows.min()))
        if t.relu_margin_rows is None or t.relu_margin_rows.shape != rows.shape:
            t.relu_margin_rows = rows.copy() if t.relu_margin_rows is None else t.relu_margin_rows
        else:
            t.relu_margin_rows = np.minimum(t.relu_margin_rows, rows)

    def bwd():
        if y.grad is not None:
            x.acc(y.grad * (x.v > 0))

    t.push(bwd)
    return y


def sigmoid(t, x):
    s = _sigmoid(x.v)
    y = Var(s)

    def bwd():
        if y.grad is not None:
            x.acc(y.grad * s * (1 - s))

    t.push(bwd)
    return y


def tanh(t, x):
    h = np.tanh(x.v)
    y = Var(h)

    def bwd():
        if y.grad is not None:
            x.acc(y.grad * (1 - h * h))

    t.push(bwd)
    return y


def mul(t, a, b):
    y = Var(a.v * b.v)

    def bwd():
        if y.grad is not None:
            a.acc(y.grad * b.v)
            b.acc(y.grad * a.v)

    t.push(bwd)
    return y


def add(t, a, b):
    y = Var(a.v + b.v)

    def bwd():
        if y.grad is not None:
            a.acc(y.grad)
            b.acc(y.grad)

    t.push(bwd)
    return y


def one_minus(t, a):
    y = Var(1 - a.v)

    def bwd():
        if y.grad is not None:
            a.acc(-y.grad)

    t.push(bwd)
    return y


def cat(t, a, b):
    y = Var(np.concatenate([a.v, b.v], axis=-1))
    n = a.v.shape[-1]

    def bwd():
        if y.grad is not None:
            a.acc(y.grad[..., :n])
            b.acc(y.grad[..., n:])

    t.push(bwd)
    return y


def dropout(t, x, p, mask):
    """nn.Dropout.  ``mask`` None => eval mode (identity); else a {0,1} keep mask."""
    if mask is None:
        return x
    scale = 1.0 / (1.0 - p)
    m = mask.astype(x.v.dtype) * x.v.dtype.type(scale)
    y = Var(x.v * m)

    def bwd():
        if y.grad is not None:
            x.acc(y.grad * m)

    t.push(bwd)
    return y


def layernorm(t, x, g, b):
    """nn.LayerNorm over the last dim, biased variance, eps=1e-5, affine."""
    mu = x.v.mean(axis=-1, keepdims=True)
    xc = x.v - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    r = 1.0 / np.sqrt(var + x.v.dtype.type(LN_EPS))
    xh = xc * r
    y = Var(xh * g.v + b.v)

    def bwd():
        if y.grad is None:
            return
        dy = y.grad
        red = tuple(range(dy.ndim - 1))
        g.acc((dy * xh).sum(axis=red))
        b.acc(dy.sum(axis=red))
        dxh = dy * g.v
        x.acc(r * (dxh - dxh.mean(axis=-1, keepdims=True) - xh * (dxh * xh).mean(axis=-1, keepdims=True)))

    t.push(bwd)
    return y


def mha(t, q, k, v, in_w, in_b, out_w, out_b, num_heads):
    """nn.MultiheadAttention(batch_first=False, dropout=0): inputs (S, B, D).

    Packed in_proj_weight = [W_q; W_k; W_v].  Written for general (S_q, S_kv); the
    reference always calls it with S_q = S_kv = 1 (multimodalIntraInterModal.py:190-197)
    where the softmax is identically 1 and W_q/W_k receive exact-zero gradients.
    """
    Sq, B, D = q.v.shape
    Sk = k.v.shape[0]
    H = num_heads
    if D % H != 0:
        raise AssertionError("embed_dim must be divisible by num_heads")
    hd = D // H
    Wq, Wk, Wv = in_w.v[:D], in_w.v[D:2 * D], in_w.v[2 * D:]
    bq, bk, bv = in_b.v[:D], in_b.v[D:2 * D], in_b.v[2 * D:]
    Q = q.v @ Wq.T + bq
    K = k.v @ Wk.T + bk
    V = v.v @ Wv.T + bv
    scale = q.v.dtype.type(1.0 / np.sqrt(hd))
    Qh = Q.reshape(Sq, B, H, hd).transpose(1, 2, 0, 3) * scale      # B,H,Sq,hd
    Kh = K.reshape(Sk, B, H, hd).transpose(1, 2, 0, 3)
    Vh = V.reshape(Sk, B, H, hd).transpose(1, 2, 0, 3)
    S = Qh @ Kh.transpose(0, 1, 3, 2)                                # B,H,Sq,Sk
    S = S - S.max(axis=-1, keepdims=True)
    P = np.exp(S)
    P = P / P.sum(axis=-1, keepdims=True)
    Oh = P @ Vh                                                      # B,H,Sq,hd
    O = Oh.transpose(2, 0, 1, 3).reshape(Sq, B, D)
    y = Var(O @ out_w.v.T + out_b.v)

    def bwd():
        if y.grad is None:
            return
        dy = y.grad
        out_w.acc(dy.reshape(-1, D).T @ O.reshape(-1, D))
        out_b.acc(dy.reshape(-1, D).sum(axis=0))
        dO = dy @ out_w.v
        dOh = dO.reshape(Sq, B, H, hd).transpose(1, 2, 0, 3)
        dVh = P.transpose(0, 1, 3, 2) @ dOh
        dP = dOh @ Vh.transpose(0, 1, 3, 2)
        dS = P * (dP - (dP * P).sum(axis=-1, keepdims=True))
        dQh = (dS @ Kh) * scale
        dKh = dS.transpose(0, 1, 3, 2) @ Qh
        dQ = dQh.transpose(2, 0, 1, 3).reshape(Sq, B, D)
        dK = dKh.transpose(2, 0, 1, 3).reshape(Sk, B, D)
        dV = dVh.transpose(2, 0, 1, 3).reshape(Sk, B, D)
        dW = np.concatenate([
            dQ.reshape(-1, D).T @ q.v.reshape(-1, D),
            dK.reshape(-1, D).T @ k.v.reshape(-1, D),
            dV.reshape(-1, D).T @ v.v.reshape(-1, D)], axis=0)
        db = np.concatenate([dQ.reshape(-1, D).sum(0), dK.reshape(-1, D).sum(0), dV.reshape(-1, D).sum(0)])
        in_w.acc(dW)
        in_b.acc(db)
        q.acc(dQ @ Wq)
        k.acc(dK @ Wk)
        v.acc(dV @ Wv)

    t.push(bwd)
    return y


def unsqueeze0(t, x):
    y = Var(x.v[None])

    def bwd():
        if y.grad is not None:
            x.acc(y.grad[0])

    t.push(bwd)
    return y


def squeeze0(t, x):
    y = Var(x.v[0])

    def bwd():
        if y.grad is not None:
            x.acc(y.grad[None])

    t.push(bwd)
    return y


# --------------------------------------------------------------------------- loss
def weighted_cross_entropy(logits, labels, class_w=None, denom=None):
    """nn.CrossEntropyLoss(weight=w, reduction='mean') forward + dlogits.

    loss = sum_i w[y_i] (lse(z_i) - z_i[y_i]) / sum_i w[y_i]   (train_pad_20.py:52,111)
    ``denom`` overrides the denominator (global sum of w[y] under data parallelism).
    Returns (loss, dlogits, numerator, denominator).
    """
    z = logits
    B, C = z.shape
    w = np.ones(C, dtype=z.dtype) if class_w is None else class_w.astype(z.dtype)
    labels = np.asarray(labels)
    valid = (labels >= 0) & (labels < C)            # nn.CrossEntropyLoss ignore_index (-100): ignored rows weigh nothing
    safe = np.where(valid, labels, 0)
    zmax = z.max(axis=1, keepdims=True)
    e = np.exp(z - zmax)
    se = e.sum(axis=1, keepdims=True)
    lse = (np.log(se) + zmax)[:, 0]
    wy = np.where(valid, w[safe], 0).astype(z.dtype)
    num = (wy * (lse - z[np.arange(B), safe])).sum()
    den = wy.sum() if denom is None else z.dtype.type(denom)
    sm = e / se
    onehot = np.zeros_like(z)
    onehot[np.arange(B), safe] = 1
    dz = (wy / den)[:, None] * (sm - onehot)
    return num / den, dz, num, wy.sum()


# --------------------------------------------------------------------------- head
class HeadConfig:
    """Shape/config of one head instance (mirrors the reference ctor arguments)."""

    def __init__(self, mechanism, F, C, V=None, T=512, D=512, H=8, n=2, text_model="one-hot-encoder"):
        self.mechanism, self.F, self.C, self.V, self.T, self.D, self.H, self.n = mechanism, F, C, V, T, D, H, n
        self.text_model = text_model

    def metablock_dims(self):
        """(V_dim, U_dim) of meta_block - multimodalIntraInterModal.py:112-115."""
        m = self.mechanism
        vdim = self.D if m == RG_ATT + "+metablock" else self.F
        udim = self.D if m in (RG_ATT + "+metablock", "metablock-se") else self.T
        return vdim, udim

    def param_shapes(self):
        """Every head parameter (reference state_dict order and names), name -> shape."""
        D, F, C, T = self.D, self.F, self.C, self.T
        s = {}

        def lin(name, o, i):
            s[name + ".weight"] = (o, i)
            s[name + ".bias"] = (o,)

        def ln(name, d):
            s[name + ".weight"] = (d,)
            s[name + ".bias"] = (d,)

        def attn(name):
            s[name + ".in_proj_weight"] = (3 * D, D)
            s[name + ".in_proj_bias"] = (3 * D,)
            lin(name + ".out_proj", D, D)

        lin("image_projector", D, F)
        if self.text_model == "one-hot-encoder":
            lin("text_fc.0", 256, self.V)
            lin("text_fc.2", 512, 256)
            lin("text_fc.4", T, 512)
        lin("text_projector", D, T)
        for a in ("image_self_attention", "text_self_attention", "image_cross_attention", "text_cross_attention"):
            attn(a)
        lin("img_gate", D, D)
        lin("txt_gate", D, D)
        vdim, udim = self.metablock_dims()
        for br in ("fb", "gb"):
            lin(f"meta_block.{br}.0", vdim, udim)
            ln(f"meta_block.{br}.1", vdim)
        for r in ("image_residual", "text_residual"):
            ln(r + ".norm", D)
            attn(r + ".attn")
            lin(r + ".gate_linear", D, D)
        nn_ = 1 if self.mechanism == "no-metadata" else self.n
        for pre, first_in, in_scope in (("fc_fusion", D * nn_, True),):
            lin(pre + ".0", D, first_in)
            ln(pre + ".1", D)
            lin(pre + ".4", D // 2, D)
            ln(pre + ".5", D // 2)
            lin(pre + ".8", C, D // 2)
        lin("fc_visual_only", C, F)
        lin("fc_fusion_proj_feat2output", C, D)
        pre = "fc_mlp_module_after_metablock_fusion_module"
        lin(pre + ".0", D, F)
        ln(pre + ".1", D)
        lin(pre + ".4", D // 2, D)
        ln(pre + ".5", D // 2)
        lin(pre + ".8", C, D // 2)
        return s


def head_forward_backward(cfg, params, img_feat, text_in, labels=None, class_w=None,
                          masks=None, denom=None, need_input_grad=False, dlogits=None):
    """Forward (+ backward when ``labels`` or ``dlogits`` is given) of the fusion head.

    params   : dict name -> ndarray (reference state_dict names, head only)
    img_feat : [B, F] backbone features (multimodalIntraInterModal.py:167-170 output)
    text_in  : [B, V] one-hot metadata, or [B, T] encoder output when cfg.text_model
               is not the one-hot encoder (the path starts at txt_feat - SURVEY §8c)
    masks    : None (eval) or dict of {0,1} keep-masks, keys among
               'img_res','txt_res' [B,D] (p=0.1), 'fc1' [B,D], 'fc2' [B,D/2] (p=0.5, or
               0.3 for the after-metablock MLP), 'img_res2','txt_res2' for the strings
               that call the residual blocks twice.
    Returns dict(logits, loss, grads{name->ndarray|None}, d_img_feat, d_text_in, num, den).
    """
    t = Tape()
    P = {k: Var(np.asarray(v), k) for k, v in params.items()}
    m = cfg.mechanism
    H = cfg.H
    mk = (lambda k: None) if masks is None else (lambda k: masks.get(k))

    x = Var(np.asarray(img_feat))
    tin = Var(np.asarray(text_in))

    def L(name, inp):
        return linear(t, inp, P[name + ".weight"], P[name + ".bias"])

    def A(name, q, k, v, heads=H):
        return mha(t, q, k, v, P[name + ".in_proj_weight"], P[name + ".in_proj_bias"],
                   P[name + ".out_proj.weight"], P[name + ".out_proj.bias"], heads)

    def residual(name, q, k, v, mask_key):
        # gatedResidualBlock.py:12-17 (heads hard-coded to 8 at :8)
        a = A(name + ".attn", q, k, v, heads=8)
        a = dropout(t, a, 0.1, None if mk(mask_key) is None else mk(mask_key)[None])
        g = sigmoid(t, L(name + ".gate_linear", q))
        out = add(t, mul(t, g, a), mul(t, one_minus(t, g), q))
        return layernorm(t, out, P[name + ".norm.weight"], P[name + ".norm.bias"])

    def mlp(pre, inp, p):
        # fc_mlp_module / fc_mlp_module_after_metablock (multimodalIntraInterModal.py:134-160)
        h = L(pre + ".0", inp)
        h = layernorm(t, h, P[pre + ".1.weight"], P[pre + ".1.bias"])
        h = dropout(t, relu(t, h), p, mk("fc1"))
        h = L(pre + ".4", h)
        h = layernorm(t, h, P[pre + ".5.weight"], P[pre + ".5.bias"])
        h = dropout(t, relu(t, h), p, mk("fc2"))
        return L(pre + ".8", h)

    def metablock(v, u):
        # metablock.py:22-32
        t1 = layernorm(t, L("meta_block.fb.0", u), P["meta_block.fb.1.weight"], P["meta_block.fb.1.bias"])
        t2 = layernorm(t, L("meta_block.gb.0", u), P["meta_block.gb.1.weight"], P["meta_block.gb.1.bias"])
        return sigmoid(t, add(t, tanh(t, mul(t, v, t1)), t2))

    # ---- common prefix (:172-200)
    p_img = L("image_projector", x)
    if cfg.text_model == "one-hot-encoder":
        h = relu(t, L("text_fc.0", tin))
        h = relu(t, L("text_fc.2", h))
        txt_feat = L("text_fc.4", h)
    else:
        txt_feat = tin
    p_txt = L("text_projector", txt_feat)
    img_seq = unsqueeze0(t, p_img)
    txt_seq = unsqueeze0(t, p_txt)
    img_att = A("image_self_attention", img_seq, img_seq, img_seq)
    txt_att = A("text_self_attention", txt_seq, txt_seq, txt_seq)
    img_cross = A("image_cross_attention", img_att, txt_att, txt_att)
    txt_cross = A("text_cross_attention", txt_att, img_att, img_att)
    img_pooled = squeeze0(t, img_cross)
    txt_pooled = squeeze0(t, txt_cross)

    def cross_pair(a, b):
        ic = A("image_cross_attention", a, b, b)
        tc = A("text_cross_attention", b, a, a)
        return ic, tc

    def gate_pair(a, b, swap=False):
        ga = sigmoid(t, L("img_gate", a))
        gb = sigmoid(t, L("txt_gate", b))
        if swap:
            ga, gb = gb, ga
        return cat(t, mul(t, ga, a), mul(t, gb, b))

    # ---- fusion strings (:205-416)
    if m == "no-metadata":
        logits = mlp("fc_fusion", p_img, 0.5)
    elif m == "no-metadata-without-mlp":
        logits = L("fc_visual_only", x)
    elif m == "concatenation":
        logits = mlp("fc_fusion", cat(t, p_img, p_txt), 0.5)
    elif m == "crossattention":
        logits = mlp("fc_fusion", cat(t, img_pooled, txt_pooled), 0.5)
    elif m == "weighted":
        logits = mlp("fc_fusion", gate_pair(p_img, p_txt), 0.5)
    elif m == "gfcam":
        logits = mlp("fc_fusion", gate_pair(img_pooled, txt_pooled), 0.5)
    elif m == "cross-weights-after-crossattention":
        logits = mlp("fc_fusion", gate_pair(img_pooled, txt_pooled, swap=True), 0.5)
    elif m == "metablock":
        logits = mlp("fc_mlp_module_after_metablock_fusion_module", metablock(x, txt_feat), 0.3)
    elif m == "rg-att2fusefeatures":
        r = squeeze0(t, residual("image_residual", txt_seq, img_seq, img_seq, "img_res"))
        logits = L("fc_fusion_proj_feat2output", r)
    elif m == "rg-att":
        ir = squeeze0(t, residual("image_residual", img_seq, txt_seq, txt_seq, "img_res"))
        tr = squeeze0(t, residual("text_residual", txt_seq, img_seq, img_seq, "txt_res"))
        logits = mlp("fc_fusion", cat(t, ir, tr), 0.5)
    elif m == "att-intramodal":
        logits = mlp("fc_fusion", cat(t, squeeze0(t, img_att), squeeze0(t, txt_att)), 0.5)
    elif m == "att-intramodal+residual":
        ir = squeeze0(t, residual("image_residual", img_seq, img_att, img_att, "img_res"))
        tr = squeeze0(t, residual("text_residual", txt_seq, txt_att, txt_att, "txt_res"))
        logits = mlp("fc_fusion", cat(t, ir, tr), 0.5)
    elif m == "cross-attention-only":
        ic, tc = cross_pair(img_seq, txt_seq)
        logits = mlp("fc_fusion", cat(t, squeeze0(t, ic), squeeze0(t, tc)), 0.5)
    elif m == "residual+cross-attention-metadados":
        ir = residual("image_residual", img_seq, img_seq, img_seq, "img_res")
        tr = residual("text_residual", txt_seq, txt_seq, txt_seq, "txt_res")
        ic, tc = cross_pair(ir, tr)
        logits = mlp("fc_fusion", cat(t, squeeze0(t, ic), squeeze0(t, tc)), 0.5)
    elif m.startswith(RG_ATT):
        ir = residual("image_residual", img_seq, img_att, img_att, "img_res")
        tr = residual("text_residual", txt_seq, txt_att, txt_att, "txt_res")
        ic, tc = cross_pair(ir, tr)
        tail = m[len(RG_ATT):]
        if tail == "":
            logits = mlp("fc_fusion", cat(t, squeeze0(t, ic), squeeze0(t, tc)), 0.5)
        elif tail == "+rg-att2fusefeatures":
            r = squeeze0(t, residual("image_residual", tc, ic, ic, "img_res2"))
            logits = L("fc_fusion_proj_feat2output", r)
        elif tail == "+metablock":
            logits = L("fc_fusion_proj_feat2output", metablock(squeeze0(t, ic), squeeze0(t, tc)))
        elif tail == "+att-intramodal+residual":
            ia = A("image_self_attention", ic, ic, ic)
            ta = A("text_self_attention", tc, tc, tc)
            ir2 = squeeze0(t, residual("image_residual", ic, ia, ia, "img_res2"))
            tr2 = squeeze0(t, residual("text_residual", tc, ta, ta, "txt_res2"))
            logits = mlp("fc_fusion", cat(t, ir2, tr2), 0.5)
        else:
            raise ValueError(f"Attention mechanism '{m}' not implemented.")
    else:
        raise ValueError(f"Attention mechanism '{m}' not implemented.")

    out = {"logits": logits.v, "loss": None, "grads": None, "d_img_feat": None, "d_text_in": None, "relu_margin": t.relu_margin,
           "relu_margin_rows": t.relu_margin_rows}
    if labels is None and dlogits is None:
        return out
    if dlogits is None:
        loss, dz, num, den = weighted_cross_entropy(logits.v, np.asarray(labels), class_w, denom)
        out.update(loss=loss, num=num, den=den, dlogits=dz)
    else:
        dz = np.asarray(dlogits)
    logits.grad = dz.astype(logits.v.dtype)
    t.backward()
    out["grads"] = {k: P[k].grad for k in params}
    if need_input_grad:
        out["d_img_feat"] = x.grad
        out["d_text_in"] = tin.grad
    return out
