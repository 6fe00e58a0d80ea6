"""CPU: the TabTransformer oracle (oracle/tabt_oracle.py) is pinned to the reference class (models/tab_transformer.py:6-60):
committed float64 golden vectors made from the unmodified class (tests/golden/tabt.npz, make_golden_tabt.py) in eval mode
and in train mode with injected dropout masks, a live re-run when /root/reference is mounted, and the host-side
contract of the fused module (state_dict names, flat parameter layout)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import tabt_oracle as to
from tests import parity
from tests.golden import make_golden_tabt as G

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tabt.npz"))


def _check(name, ref):
    cards, ncont, D, H, L, F, O, B, train = G.CASES[name]
    params, x_cat, x_num, dout, masks = G.case_inputs(name)
    r = to.forward_backward(params, x_cat, x_num, H, 0.3, masks, dout)
    assert parity.rel_err(r["out"], ref(name + "/out")) < 1e-12
    for k, g in r["grads"].items():
        assert parity.rel_err(g, ref(name + "/grad/" + k)) < 1e-11, k
    if ncont > 0:
        assert parity.rel_err(r["d_num"], ref(name + "/d_num")) < 1e-12


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_matches_golden(name):
    _check(name, lambda k: GOLD[k])


@pytest.mark.skipif(not G.reference_available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("name", ["eval_small", "train_small"])
def test_oracle_matches_live_reference(name):
    live = G.run_reference(name)
    for k, v in live.items():
        assert np.array_equal(v, GOLD[name + "/" + k]) or parity.rel_err(v, GOLD[name + "/" + k]) < 1e-13, k
    _check(name, lambda k: live[k.split("/", 1)[1]])


def test_fused_module_keeps_the_reference_state_dict_and_flat_layout():
    import fusion_b200 as fb
    from fusion_b200 import _lib
    cards = [10] * 82
    m = fb.TabTransformer(cards, num_continuous=4, output_dim=85)         # loadImageModelClassifier.py:190-198
    shapes = to.param_shapes(cards, 4, 32, 128, 2, 85)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in shapes.items()]
    d = _lib.TabtDesc(B=4, T=82, D=32, H=4, F=128, L=2, n_emb_rows=820, train=0, p=0.3, flags=0)
    layer, total = C.c_int64(), C.c_int64()
    _lib.check(_lib.lib().fb200_tabt_param_elems(C.byref(d), C.byref(layer), C.byref(total)))
    assert layer.value == sum(int(np.prod(s)) for s in to.layer_shapes(32, 128).values())
    flat = m._flat_params()
    assert flat.numel() == total.value == 2 * layer.value + 820 * 32
    # order: layer blocks in state_dict order, then the tables
    sd = m.state_dict()
    ref = torch.cat([sd[f"transformer_encoder.layers.{l}.{k}"].reshape(-1) for l in range(2) for k in to.LAYER_KEYS] +
                    [sd[f"embeddings.{i}.weight"].reshape(-1) for i in range(82)])
    assert torch.equal(flat.detach(), ref)
    assert m._emb_base.tolist() == list(range(0, 820, 10))


def test_descriptor_checks_on_the_host():
    from fusion_b200 import _lib
    L = _lib.lib()
    ok = dict(B=4, T=82, D=32, H=4, F=128, L=2, n_emb_rows=820, train=1, p=0.3, flags=0)
    def rc(**kw):
        d = _lib.TabtDesc(**{**ok, **kw})
        n = C.c_int64()
        return L.fb200_tabt_param_elems(C.byref(d), None, C.byref(n))
    assert rc() == 0
    assert rc(H=5) == -1                      # embed_dim % num_heads (torch asserts)
    assert rc(D=30, H=3) == -2                # D % 4
    assert rc(T=400) == -2                    # one sample no longer fits one SM's shared memory
    assert rc(p=1.0) == -1
    with pytest.raises(_lib.Fb200Error):      # host tensors: there is no CPU path
        import fusion_b200 as fb
        m = fb.TabTransformer([3, 4], 2, embed_dim=8, num_heads=2, hidden_dim=16, output_dim=3).eval()
        m(torch.zeros(2, 2, dtype=torch.int64), torch.zeros(2, 2))
