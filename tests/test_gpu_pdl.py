"""GPU: programmatic dependent launch must be invisible.  Every kernel of the library is launched with the
programmatic-stream-serialization attribute and gates its first global-memory access on griddepcontrol.wait
(csrc/common.cuh: pdl_sync).  A missing or misplaced wait shows up as a reader seeing stale data now and then:
the same forward + backward is repeated many times with the attribute on and off and every repetition has to
reproduce the first one up to the order of the fp32 gradient atomics (< 1e-6 observed; a stale read is O(1)),
and the fused Adam chained between torch kernels has to be bit-identical with and without it."""
import pytest
import torch

import fusion_b200 as fb
from fusion_b200 import _lib
from tests.golden import cases as C
from tests.gpu_util import build_model, case_inputs

pytestmark = pytest.mark.gpu


def _grads_once(model, x, tin, y, cw, fused):
    model.zero_grad(set_to_none=True)
    if fused:
        loss, _ = model.forward_loss(x, tin, y, cw)
    else:
        loss = fb.FusedCrossEntropyLoss(weight=cw)(model(x, tin), y)
        loss.backward()
    return loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mech,B,dims", [
    ("crossattention", 32, C.SMALL_DIMS),                                        # FFMA path: split-K fix-up, bias-gradient atomics
    ("att-intramodal+residual+cross-attention-metadados", 32, C.SMALL_DIMS),
    ("crossattention", 256, dict(F=512, V=85, C=6)),                             # tcgen05 path, grouped weight gradients
])
def test_repeated_steps_agree_with_pdl(mech, B, dims, fused):
    L = _lib.lib()
    case = dict(cfg=dict(dims, mechanism=mech), B=B, seed=5, train=False, full_grads=False)
    cfg, model = build_model(case, "fp32")
    x, tin, y, cw, _ = case_inputs(cfg, case)
    model.eval()
    prev = L.fb200_debug_set_pdl(0)
    try:
        l_ref, g_ref = _grads_once(model, x, tin, y, cw, fused)
        L.fb200_debug_set_pdl(1)
        for _ in range(40):
            l, g = _grads_once(model, x, tin, y, cw, fused)
            assert abs(float(l) - float(l_ref)) < 2e-6 * abs(float(l_ref))
            assert g.keys() == g_ref.keys()
            for k in g_ref:
                dev = float((g[k] - g_ref[k]).abs().max() / g_ref[k].abs().max().clamp_min(1e-30))
                assert dev < 1e-5, (k, dev)
    finally:
        L.fb200_debug_set_pdl(prev)


def test_fused_adam_bitwise_with_and_without_pdl():
    L = _lib.lib()
    shapes = [(512, 2048), (512,), (6, 256), (3,), (1536, 512), (7, 13), (256, 85), (256,)]
    res = {}
    prev = L.fb200_debug_set_pdl(0)
    try:
        for pdl in (0, 1):
            L.fb200_debug_set_pdl(pdl)
            torch.manual_seed(1)
            ps = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
            opt = fb.FusedAdam(ps, lr=5e-3, weight_decay=1e-2)
            for _ in range(30):
                for p in ps:
                    p.grad = torch.randn_like(p)            # torch kernel -> fused Adam -> torch kernel -> ...
                opt.step()
            torch.cuda.synchronize()
            res[pdl] = [p.detach().clone() for p in ps]
    finally:
        L.fb200_debug_set_pdl(prev)
    assert all(torch.equal(a, b) for a, b in zip(res[0], res[1]))


@pytest.mark.parametrize("mech", ["crossattention", "gfcam", "metablock", "att-intramodal+residual+cross-attention-metadados",
                                  "att-intramodal+residual+cross-attention-metadados+att-intramodal+residual", "rg-att"])
@pytest.mark.parametrize("B", [32, 640])
def test_two_lane_execution_equals_one_stream(mech, B):
    """The metadata chain is launched on an internal side stream (plan.cu: lanes, exec.cu: LaneSync) - FFMA path with
    per-lane split-K scratch at B = 32, tcgen05 path at B = 640.
    Repeated steps must reproduce the single-stream result (FB200_FLAG_ONE_STREAM) - a missing cross-lane
    dependency shows up as a stale or half-written operand now and then."""
    dims = dict(F=512, V=85, C=6)
    case = dict(cfg=dict(dims, mechanism=mech), B=B, seed=11, train=False, full_grads=False)
    cfg, one = build_model(case, "fp32", flags=_lib.FLAG_ONE_STREAM | _lib.FLAG_NO_MEGA)
    _, two = build_model(case, "fp32", flags=_lib.FLAG_NO_MEGA)
    _, mega = build_model(case, "fp32", flags=_lib.FLAG_FORCE_MEGA)            # B = 32: the persistent step kernel (repeated steps: its grid barrier and stage order)
    x, tin, y, cw, _ = case_inputs(cfg, case)
    one.eval(); two.eval(); mega.eval()
    l_ref, g_ref = _grads_once(one, x, tin, y, cw, True)
    for model, fused in ((two, True), (two, False)) + (((mega, True), (mega, False)) if B <= 64 else ()):
        for _ in range(15):
            l, g = _grads_once(model, x, tin, y, cw, fused)
            assert abs(float(l) - float(l_ref)) < 2e-6 * abs(float(l_ref))
            assert g.keys() == g_ref.keys()
            for k in g_ref:
                dev = float((g[k] - g_ref[k]).abs().max() / g_ref[k].abs().max().clamp_min(1e-30))
                assert dev < 1e-5, (k, dev)


@pytest.mark.parametrize("graph", [False, True])
def test_dp_mid_event_marks_first_bucket_final(graph):
    """fb200_head_train_step_dp: when the mid-step event fires, every gradient below the bucket split already has its final
    value (a communication stream that waits for the event alone may all-reduce it), and the step as a whole is unchanged."""
    dims = dict(F=2048, V=85, C=6)
    case = dict(cfg=dict(dims, mechanism="crossattention"), B=1024, seed=3, train=False, full_grads=False)
    cfg, model = build_model(case, "fp32")
    x, tin, y, cw, _ = case_inputs(cfg, case)
    model.eval()
    l_ref, _ = model.forward_loss(x, tin, y, cw)
    ref = model.flat_grad.clone()
    split = _lib.dp_bucket_split(model.last_desc)
    total = ref.numel()
    assert 0 < split < total
    lo, hi = fb.dp.BucketedAllReduce.split_ranges(_lib.grad_live_ranges(model.last_desc), split)
    assert lo and hi and all(e <= split for _, e in lo) and all(b >= split for b, _ in hi)
    mid = torch.cuda.Event(); mid.record()
    side = torch.cuda.Stream()
    snap = torch.empty(split, device="cuda")
    for _ in range(5):
        if graph:
            step = fb.GraphedTrainStep(model, x, tin, y, cw, warmup=1, mid_event=mid)
            step.flat_grad.zero_()
            step.run(); flat = step.flat_grad
        else:
            model.forward_loss(x, tin, y, cw, mid_event=mid); flat = model.flat_grad
        side.wait_event(mid)                                  # NOT the end of the step
        with torch.cuda.stream(side):
            snap.copy_(flat[:split])
        torch.cuda.synchronize()
        dev = float((snap - ref[:split]).abs().max() / ref[:split].abs().max())
        assert dev < 1e-5, dev
        assert float((flat - ref).abs().max() / ref.abs().max()) < 1e-5
