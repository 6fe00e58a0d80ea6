"""CPU: the oracle's general (S_q, S_kv) multi-head attention is pinned to torch.nn.MultiheadAttention - the module the
reference instantiates (models/multimodalIntraInterModal.py:78-100; called with token sequences in
models/multimodalGated.py:118-206) - through committed float64 golden vectors (tests/golden/mha.npz, made by
tests/golden/make_golden_mha.py) and through a live comparison on fresh shapes."""
import os

import numpy as np
import pytest

from oracle import head_oracle as ho
from tests import parity
from tests.golden import make_golden_mha as G

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "mha.npz"))


def oracle_mha(c, H, self_attn):
    """c: dict with q, k, v, weights, dy -> dict of outputs / gradients from the numpy oracle."""
    t = ho.Tape()
    q = ho.Var(c["q"])
    k, v = (q, q) if self_attn else (ho.Var(c["k"]), ho.Var(c["v"]))
    in_w, in_b, out_w, out_b = (ho.Var(c[n]) for n in ("in_w", "in_b", "out_w", "out_b"))
    y = ho.mha(t, q, k, v, in_w, in_b, out_w, out_b, H)
    y.grad = c["dy"].copy()
    t.backward()
    r = dict(out=y.v, dq=q.grad, d_in_w=in_w.grad, d_in_b=in_b.grad, d_out_w=out_w.grad, d_out_b=out_b.grad)
    if not self_attn:
        r.update(dk=k.grad, dv=v.grad)
    return r


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_matches_golden(name):
    Sq, Sk, B, D, H, sa = G.CASES[name]
    c = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    got = oracle_mha(c, H, sa)
    for k, v in got.items():
        assert parity.rel_err(v, c[k]) < 1e-12, k


@pytest.mark.parametrize("shape", [(3, 11, 2, 48, 6, False), (40, 9, 1, 32, 1, False), (6, 6, 5, 80, 8, True)])
def test_oracle_matches_live_torch(shape):
    Sq, Sk, B, D, H, sa = shape
    c = G.run_torch(Sq, Sk, B, D, H, sa, seed=7)
    got = oracle_mha(c, H, sa)
    for k, v in got.items():
        assert parity.rel_err(v, c[k]) < 1e-12, k


def test_s1_query_and_key_weights_get_zero_gradient():
    # the fact the head lowering of S = 1 attention rests on (SURVEY "Facts"): softmax over one key is exactly 1
    c = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith("s1/")}
    D = c["q"].shape[-1]
    assert np.all(c["d_in_w"][: 2 * D] == 0) and np.all(c["d_in_b"][: 2 * D] == 0) and np.all(c["dq"] == 0)
