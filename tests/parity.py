"""Shared comparison helpers: an implementation's outputs vs a golden fixture / the oracle."""
from __future__ import annotations

import os

import numpy as np

from tests.golden import cases as C

FP32_TOL = 1e-5      # north_star: logits and gradients within 1e-5 relative in fp32
BF16_TOL = 2e-2      # ... and within 2e-2 relative in bf16


def rel_err(a, b):
    """max|a-b| / max|b| (tensor-wise relative error; element-wise ratios are meaningless near zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.abs(b).max() if b.size else 0.0
    d = np.abs(a - b).max() if b.size else 0.0
    return d / s if s > 0 else d


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 over the whole tensor (reported next to the max-norm figure: DESIGN.md section 5)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    n = np.linalg.norm(b)
    d = np.linalg.norm(a - b)
    return d / n if n > 0 else d


def load_golden(name):
    return np.load(os.path.join(C.GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def check_against_golden(name, case, logits, loss, grads, d_img_feat, tol, check_argmax=True):
    """grads: dict name -> ndarray or None.  Returns the worst relative error seen."""
    g = load_golden(name)
    worst = rel_err(logits, g["logits64"])
    assert worst <= tol, f"{name}: logits rel err {worst:.3e} > {tol}"
    if check_argmax:
        ref = np.asarray(g["logits64"])
        if ref.shape[1] > 1:
            srt = np.sort(ref, axis=1)
            decided = (srt[:, -1] - srt[:, -2]) > 2 * tol * np.abs(ref).max()     # rows that are not numerical ties
        else:
            decided = np.ones(ref.shape[0], bool)
        assert (np.argmax(logits, 1)[decided] == np.argmax(ref, 1)[decided]).all(), f"{name}: argmax differs"
    e = abs(float(loss) - float(g["loss64"])) / abs(float(g["loss64"]))
    assert e <= tol, f"{name}: loss rel err {e:.3e}"
    worst = max(worst, e)
    none_names = set(g["none_grads"].tolist())
    live_names = set(g["grad_names"].tolist())
    for k in none_names:
        assert grads.get(k) is None, f"{name}: {k} must keep grad=None (reference pattern)"
    items = [(k, grads.get(k)) for k in sorted(live_names)]
    if d_img_feat is not None:
        items.append(("d_img_feat", d_img_feat))
    for k, got in items:
        assert got is not None, f"{name}: {k} must receive a gradient"
        got = np.asarray(got, dtype=np.float64)
        if "g:" + k in g.files:
            e = rel_err(got, g["g:" + k])
            assert e <= tol, f"{name}: grad {k} rel err {e:.3e} > {tol}"
        else:
            maxabs, l2 = g["m:" + k]
            flat = got.ravel()
            samp = flat[g["i:" + k]]
            e = np.abs(samp - g["s:" + k]).max() / maxabs if maxabs > 0 else np.abs(samp).max()
            assert e <= tol, f"{name}: grad {k} sampled rel err {e:.3e} > {tol}"
            s = C.grad_summary(k, got, case["seed"])
            pe = np.abs(s["probes"] - g["p:" + k]).max() / l2 if l2 > 0 else np.abs(s["probes"]).max()
            assert pe <= 8 * tol, f"{name}: grad {k} projection err {pe:.3e} > {8 * tol}"
            me = abs(s["maxabs"] - maxabs) / maxabs if maxabs > 0 else s["maxabs"]
            assert me <= 4 * tol, f"{name}: grad {k} max-abs err {me:.3e}"
            e = max(e, pe / 8)
        worst = max(worst, e)
    # exact zeros where the reference has them (W_q / W_k rows of S=1 attention)
    return worst
