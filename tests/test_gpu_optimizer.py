"""GPU: fused Adam against torch.optim.Adam, and a multi-step training trajectory of the whole hot path
(fused head + fused CE + fused Adam) against the numpy oracle + numpy Adam."""
import numpy as np
import pytest
import torch

import fusion_b200 as fb
from fusion_b200 import _lib
from oracle import head_oracle as ho
from oracle.adam_oracle import AdamOracle
from tests import parity
from tests.golden import cases as C
from tests.gpu_util import build_model, case_inputs

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam():
    torch.manual_seed(0)
    shapes = [(512, 2048), (512,), (6, 256), (3,), (1536, 512), (7, 13)]
    ref = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    a = torch.optim.Adam(ref, lr=5e-3, weight_decay=1e-2)
    b = fb.FusedAdam(mine, lr=5e-3, weight_decay=1e-2)
    for it in range(7):
        for i, (p, q) in enumerate(zip(ref, mine)):
            if i == 3 and it < 3:          # a parameter without gradient for the first steps (None-grad skip)
                p.grad = q.grad = None
                continue
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        a.step(); b.step()
    for p, q in zip(ref, mine):
        assert parity.rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) < 2e-6
    sa, sb = a.state_dict()["state"], b.state_dict()["state"]
    assert set(sa[0].keys()) == {"step", "exp_avg", "exp_avg_sq"} <= set(sb[0].keys()) | {"step"}
    assert int(sb[3]["step"]) == 4


def test_training_trajectory_matches_oracle():
    case = dict(cfg=dict(C.SMALL_DIMS, mechanism="crossattention"), B=32, seed=31, train=False, full_grads=False)
    cfg, model = build_model(case, "fp32")
    x, tin, y, cw, _ = case_inputs(cfg, case)
    model.eval()                                            # no dropout: the trajectory is deterministic
    opt = fb.FusedAdam(model.parameters(), lr=5e-3, weight_decay=1e-4)
    crit = fb.FusedCrossEntropyLoss(weight=cw)
    params = C.gen_params(cfg, case["seed"], np.float64)
    xn, tn, labels, cwn, _ = C.gen_inputs(cfg, case["B"], case["seed"], False, np.float64)
    oracle_opt = AdamOracle(params, lr=5e-3, weight_decay=1e-4)
    losses, ref_losses = [], []
    for step in range(25):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x, tin), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
        o = ho.head_forward_backward(cfg, params, xn, tn, labels, cwn, None)
        ref_losses.append(float(o["loss"]))
        oracle_opt.step(o["grads"])
    assert ref_losses[-1] < 0.7 * ref_losses[0]             # it actually trains
    dev = np.abs(np.array(losses) - np.array(ref_losses)) / np.array(ref_losses)
    # The first steps are well conditioned: fp32 kernels against the float64 oracle.  Further on, Adam turns the
    # 1e-7 noise of the fp32 gradient atomics (their order varies from run to run) into lr-sized differences
    # whenever a ReLU pre-activation sits within that noise of zero, so the tail is held to a looser bound.
    assert np.max(dev[:8]) < 2e-5
    assert np.max(dev) < 2e-3
    sd = model.state_dict()
    for k in ("image_projector.weight", "fc_fusion.8.weight", "text_fc.0.bias"):
        assert parity.rel_err(sd[k].cpu().numpy(), params[k]) < 5e-3, k
    # parameters the mechanism never touches were not decayed (grad None => skipped, as in the reference)
    untouched = C.gen_params(cfg, case["seed"], np.float32)["img_gate.weight"]
    assert np.array_equal(sd["img_gate.weight"].cpu().numpy(), untouched)


def test_graphed_train_step_equals_eager():
    case = dict(cfg=dict(mechanism="crossattention", F=2048, V=85, C=6), B=512, seed=9, train=True, full_grads=False)
    cfg, model = build_model(case, "fp32")
    x, tin, y, cw, _ = case_inputs(cfg, case)
    model.train()
    model._rng_state = None
    loss0, logits0 = model.forward_loss(x, tin, y, cw)
    g0 = model.flat_grad.clone(); l0 = float(loss0)
    model._rng_state[1] = 1                                   # rewind the Philox offset: the graph must redraw the same masks
    step = fb.GraphedTrainStep(model, x, tin, y, cw, warmup=0)
    model._rng_state[1] = 1
    step.run(); torch.cuda.synchronize()
    assert abs(float(step.loss) - l0) < 1e-6 * abs(l0)
    assert parity.rel_err(step.flat_grad.cpu().numpy(), g0.cpu().numpy()) < 1e-5
    a = float(step.run()); b = float(step.run())
    assert a != b                                            # fresh dropout masks on every replay (device-side Philox offset)


def test_graphed_small_batch_captures_the_tensor_core_engine():
    """B = 32 (the reference's BATCH_SIZE): an eager call is ONE launch of the persistent step kernel, a captured step takes the
    per-op tcgen05 kernels with cluster split-K (faster inside a graph) - same loss and gradients within the fp32 bar, and the
    model's own engine flags are left as they were."""
    case = dict(cfg=dict(mechanism="crossattention", F=2048, V=85, C=6), B=32, seed=4, train=False, full_grads=False)
    cfg, model = build_model(case, "fp32")
    x, tin, y, cw, _ = case_inputs(cfg, case)
    model.eval()
    loss0, _ = model.forward_loss(x, tin, y, cw)
    assert _lib.mega_program_info(model.last_desc) is not None          # eager: the step kernel
    g0 = model.flat_grad.clone(); l0 = float(loss0)
    step = fb.GraphedTrainStep(model, x, tin, y, cw, warmup=1)
    assert model.engine_flags == 0 and (step.desc.flags & _lib.FLAG_FORCE_TC)
    assert _lib.mega_program_info(step.desc) is None                    # captured: per-op kernels
    step.run(); torch.cuda.synchronize()
    assert abs(float(step.loss) - l0) < 1e-5 * abs(l0)
    assert parity.rel_err(step.flat_grad.cpu().numpy(), g0.cpu().numpy()) < 1e-5
    forced = fb.GraphedTrainStep(build_model(case, "fp32", flags=_lib.FLAG_FORCE_MEGA)[1].eval(), x, tin, y, cw, warmup=1)
    assert _lib.mega_program_info(forced.desc) is not None              # an engine the caller chose is kept
