"""CPU: the timing port (oracle/torch_port.py, what `--impl reference` / cpu_baseline run on the GPU
box) is the reference: bit-identical outputs and gradients under the same seed, whenever the
reference sources are present; and it agrees with the numpy oracle everywhere."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from oracle import ref_shim
from oracle.torch_port import ReferencePort, time_cpu_train_step
from tests import parity
from tests.golden import cases as C


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present")
@pytest.mark.parametrize("mech", ["concatenation", "crossattention", "metablock", "weighted", "gfcam", ho.RG_ATT, ho.RG_ATT + "+metablock"])
def test_port_is_bit_identical_to_reference(mech):
    torch.manual_seed(3)
    ref = ref_shim.build_reference_model(mech, 256, 6, V=13, D=64).eval()
    torch.manual_seed(3)
    port = ReferencePort(mech, 256, 6, V=13, D=64).eval()
    port.load_state_dict(ref.state_dict(), strict=True)
    x, t = torch.randn(9, 256), torch.randn(9, 13)
    y = torch.randint(0, 6, (9,))
    la, lb = ref(x, t), port(x, t)
    assert torch.equal(la, lb)
    torch.nn.functional.cross_entropy(la, y).backward()
    torch.nn.functional.cross_entropy(lb, y).backward()
    for (k, p), (_, q) in zip(ref.named_parameters(), port.named_parameters()):
        assert (p.grad is None) == (q.grad is None), k
        if p.grad is not None:
            assert torch.equal(p.grad, q.grad), k


@pytest.mark.parametrize("name", ["small02_train", "small07_train", "small14_train"])
def test_port_matches_golden(name):
    case = C.all_cases()[name]
    cfg = C.make_cfg(case["cfg"])
    port = ReferencePort(cfg.mechanism, cfg.F, cfg.C, V=cfg.V, T=cfg.T, D=cfg.D, H=cfg.H).double().eval()
    params = C.gen_params(cfg, case["seed"], np.float64)
    port.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    x, tin, labels, cw, _ = C.gen_inputs(cfg, case["B"], case["seed"], False, np.float64)
    logits = port(torch.from_numpy(x), torch.from_numpy(tin)).detach().numpy()
    o = ho.head_forward_backward(cfg, params, x, tin)
    assert parity.rel_err(logits, o["logits"]) < 1e-12


def test_cpu_timer_runs():
    r = time_cpu_train_step("concatenation", 64, 13, 6, 8, D=64, steps=2, warmup=1, threads=1, budget_s=5)
    assert r["samples_per_s"] > 0 and r["steps"] >= 2
