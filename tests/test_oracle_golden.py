"""CPU: the numpy oracle reproduces every golden fixture (generated from the UNMODIFIED
reference by tests/golden/make_golden.py) - logits, loss, every gradient, None pattern."""
import numpy as np
import pytest

from oracle import head_oracle as ho
from tests import parity
from tests.golden import cases as C

CASES = C.all_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    case = CASES[name]
    cfg = C.make_cfg(case["cfg"])
    params = C.gen_params(cfg, case["seed"], np.float64)
    x, tin, labels, cw, masks = C.gen_inputs(cfg, case["B"], case["seed"], case["train"], np.float64)
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    dx = o["d_img_feat"] if o["d_img_feat"] is not None else np.zeros_like(x)
    worst = parity.check_against_golden(name, case, o["logits"], o["loss"], o["grads"], dx, tol=2e-7)   # full gradients are stored as float32
    assert worst < 2e-7


def test_qk_rows_are_exact_zeros():
    """S=1 attention: W_q / W_k (first 2D rows of in_proj) get materialised exact zeros."""
    case = CASES["small03_train"]
    cfg = C.make_cfg(case["cfg"])
    params = C.gen_params(cfg, case["seed"], np.float64)
    x, tin, labels, cw, masks = C.gen_inputs(cfg, case["B"], case["seed"], True, np.float64)
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks)
    for a in ("image_self_attention", "text_self_attention", "image_cross_attention", "text_cross_attention"):
        g = o["grads"][a + ".in_proj_weight"]
        assert (g[: 2 * cfg.D] == 0).all() and np.abs(g[2 * cfg.D:]).max() > 0
        assert (o["grads"][a + ".in_proj_bias"][: 2 * cfg.D] == 0).all()


def test_num_heads_is_irrelevant_at_s1():
    case = CASES["small03_train"]
    kw = dict(case["cfg"])
    outs = []
    for H in (2, 4, 8):
        kw["H"] = H
        cfg = C.make_cfg(kw)
        params = C.gen_params(cfg, case["seed"], np.float64)
        x, tin, labels, cw, masks = C.gen_inputs(cfg, case["B"], case["seed"], True, np.float64)
        outs.append(ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks)["logits"])
    assert np.abs(outs[0] - outs[1]).max() < 1e-12 and np.abs(outs[0] - outs[2]).max() < 1e-12


def test_weighted_ce_matches_definition():
    rng = np.random.default_rng(0)
    z = rng.standard_normal((7, 5))
    y = rng.integers(0, 5, 7)
    w = rng.random(5) + 0.5
    loss, dz, num, den = ho.weighted_cross_entropy(z, y, w)
    lse = np.log(np.exp(z).sum(1))
    ref = (w[y] * (lse - z[np.arange(7), y])).sum() / w[y].sum()
    assert abs(loss - ref) < 1e-12
    eps = 1e-6
    zp = z.copy(); zp[2, 3] += eps
    lp = ho.weighted_cross_entropy(zp, y, w)[0]
    assert abs((lp - loss) / eps - dz[2, 3]) < 1e-5
    # global denominator (data parallel): the two halves' gradients add up to the full batch's
    _, d1, n1, w1 = ho.weighted_cross_entropy(z[:4], y[:4], w, denom=den)
    _, d2, n2, w2 = ho.weighted_cross_entropy(z[4:], y[4:], w, denom=den)
    assert np.abs(np.concatenate([d1, d2]) - dz).max() < 1e-12 and abs((n1 + n2) / (w1 + w2) - loss) < 1e-12
