"""CPU, build container only: pin the numpy oracle against the UNMODIFIED reference module
imported from /root/reference (skipped on the GPU box, where the reference is absent)."""
import numpy as np
import pytest

from oracle import head_oracle as ho
from oracle import ref_shim
from tests import parity
from tests.golden import cases as C

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present")
CASES = C.all_cases()
PICK = ["cfg2_cross_train", "cfg3a_meta_train", "cfg5_rgatt_train", "small06_train", "small15_train", "small16_train", "small17_train", "edge_cross_B1"]


@pytest.mark.parametrize("name", PICK)
def test_live_reference(name):
    import torch
    from tests.golden.make_golden import run_reference
    case = CASES[name]
    cfg = C.make_cfg(case["cfg"])
    logits, loss, grads, dx = run_reference(case, np.float64)
    params = C.gen_params(cfg, case["seed"], np.float64)
    x, tin, labels, cw, masks = C.gen_inputs(cfg, case["B"], case["seed"], case["train"], np.float64)
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    assert parity.rel_err(o["logits"], logits) < 1e-12
    assert abs(o["loss"] - loss) < 1e-12
    for k, g in grads.items():
        if g is None:
            assert o["grads"][k] is None, k
        else:
            assert parity.rel_err(o["grads"][k], g) < 1e-11, k
    assert parity.rel_err(o["d_img_feat"], dx) < 1e-11
