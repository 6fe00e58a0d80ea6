"""bf16-rounding oracle (TEST INFRASTRUCTURE): the float64 numpy oracle with the CUDA bf16 path's rounding points.

The CUDA bf16 mode keeps fp32 master weights and fp32 accumulation and rounds to bf16 exactly
  * every internal activation buffer that some Linear reads (the value is STORED as bf16, so every later consumer -
    gate, gated residual, the ReLU mask of the backward pass - sees the rounded value),
  * the external fp32 inputs, once, for the tensor-core Linears that read them (other consumers keep the fp32 input),
  * the weights of the tensor-core Linears (widths multiple of 8 and >= 32; K = 85 / 13 / 11 and the C-wide classifier
    run on the fp32 FFMA kernels with fp32 weights),
  * the gradient buffer of every Linear output (dY is stored as bf16 and read by dX, dW and db alike), except dlogits.
`emulate(cfg)` patches oracle.head_oracle.linear / .mha (S = 1: out_proj(v_proj(kv)), what the CUDA plan lowers to)
accordingly for the duration of a `with` block.  Everything else stays float64: what remains between this oracle and
the CUDA result is fp32-vs-float64 accumulation and the order of the partial roundings of accumulated gradients.
"""
from __future__ import annotations

import contextlib

import numpy as np

from oracle import head_oracle as ho


def bf16(a):
    """Round-to-nearest-even to bfloat16, returned as float64."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).astype(np.float64)


def _tc_shape(K, N):
    return K % 8 == 0 and N % 8 == 0 and K >= 32 and N >= 32           # plan.cu: Builder::linear engine choice


@contextlib.contextmanager
def emulate(cfg, ext_inputs):
    """ext_inputs: the numpy arrays handed to head_forward_backward as img_feat / text_in (identified by memory)."""
    ext_ids = [np.asarray(a) for a in ext_inputs]

    def is_ext(v):
        return any(np.shares_memory(v, e) for e in ext_ids)

    def linear_q(t, x, W, b):
        N, K = W.v.shape
        tc = _tc_shape(K, N)
        ext = is_ext(x.v)
        if ext:
            xq = bf16(x.v) if tc else x.v
        else:
            xq = bf16(x.v)
            x.v[...] = xq                                 # stored as bf16: later consumers see the rounded value
        Wq = bf16(W.v) if tc else W.v
        y = ho.Var(xq @ Wq.T + b.v)
        is_logits = (N == cfg.C)

        def bwd():
            if y.grad is None:
                return
            dyq = y.grad if is_logits else bf16(y.grad)
            x.acc(dyq @ Wq)
            g2 = dyq.reshape(-1, dyq.shape[-1])
            W.acc(g2.T @ xq.reshape(-1, xq.shape[-1]))
            b.acc(g2.sum(axis=0))

        t.push(bwd)
        return y

    def mha_s1(t, q, k, v, in_w, in_b, out_w, out_b, num_heads):
        Sq, B, D = q.v.shape
        assert Sq == 1 and k.v.shape[0] == 1, "the bf16 emulation covers the S = 1 attention of the head"
        if D % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        Wv, bv = bf16(in_w.v[2 * D:]), in_b.v[2 * D:]
        Wo = bf16(out_w.v)
        if not is_ext(v.v):
            v.v[...] = bf16(v.v)
        vq = bf16(v.v)
        A = bf16(vq @ Wv.T + bv)                          # stored as bf16 (the out-projection reads it)
        y = ho.Var(A @ Wo.T + out_b.v)

        def bwd():
            if y.grad is None:
                return
            dy = bf16(y.grad).reshape(-1, D)
            out_w.acc(dy.T @ A.reshape(-1, D))
            out_b.acc(dy.sum(axis=0))
            dA = bf16(dy @ Wo)
            dW = np.zeros_like(in_w.v); db = np.zeros_like(in_b.v)
            dW[2 * D:] = dA.T @ vq.reshape(-1, D)
            db[2 * D:] = dA.sum(axis=0)
            in_w.acc(dW); in_b.acc(db)
            v.acc((dA @ Wv).reshape(v.v.shape))
            q.acc(np.zeros_like(q.v)); 
            if k is not v:
                k.acc(np.zeros_like(k.v))

        t.push(bwd)
        return y

    old = ho.linear, ho.mha
    ho.linear, ho.mha = linear_q, mha_s1
    try:
        yield
    finally:
        ho.linear, ho.mha = old


def forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True):
    """head_forward_backward under the bf16 rounding points (inputs are copied: the emulation rounds buffers in place)."""
    x = np.array(x, dtype=np.float64); tin = np.array(tin, dtype=np.float64)
    params = {k: np.array(v, dtype=np.float64) for k, v in params.items()}
    with emulate(cfg, (x, tin)):
        return ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=need_input_grad)
