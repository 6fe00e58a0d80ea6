"""GPU: what the fifth session of round 2 added - the column-owner backward of the classifier head (one and two columns per
thread, small / ragged batches, narrow heads) against the float64 oracle, and the grid-wide GEMM timeline debug facility."""
import ctypes as Ct

import numpy as np
import pytest
import torch

from fusion_b200 import _lib
from oracle import head_oracle as ho
from tests import parity
from tests.golden import cases as C
from tests.gpu_util import build_model, run_autograd_arrays, tie_free_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mech,D,Cn,B", [("crossattention", 1024, 6, 96),    # classifier K = 512: two columns per thread
                                          ("crossattention", 1024, 8, 41),    # ragged rows, all eight class slots
                                          ("concatenation", 64, 2, 300),      # K = 32: most threads own no column
                                          ("metablock", 512, 8, 7)])          # fewer rows than one trip of four... per CTA
def test_classifier_backward_columns(mech, D, Cn, B):
    """smalln_bwd_cols_kernel through the per-op engines (FFMA GEMMs so that every batch takes the stand-alone launch, not the
    persistent step kernel): dX, dW, db of the class head and everything upstream of it against the oracle, 1e-5 relative."""
    kw = dict(mechanism=mech, F=256, V=13, C=Cn, D=D, H=8)
    case = dict(cfg=kw, B=B, seed=3100 + D + B, train=True, full_grads=False)
    cfg = C.make_cfg(kw)
    params = C.gen_params(cfg, case["seed"], np.float64)
    x, tin, labels, cw, masks = tie_free_inputs(cfg, params, B, case["seed"])
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    for flags in (_lib.FLAG_FORCE_SIMT, _lib.FLAG_FORCE_TC):
        cfg, model = build_model(case, "fp32", flags=flags)
        logits, loss, grads, dx = run_autograd_arrays(model, x, tin, labels, cw, masks)
        assert parity.rel_err(logits, o["logits"]) < parity.FP32_TOL
        assert abs(loss - o["loss"]) < parity.FP32_TOL * abs(o["loss"])
        for k, g in o["grads"].items():
            if g is None:
                assert grads[k] is None, k
            else:
                assert parity.rel_err(grads[k], g) < parity.FP32_TOL, (k, flags)


def test_gemm_timeline_records_every_cta():
    """fb200_debug_tc_timeline: one 1024 x 512 x 512 fp32-strict GEMM = 8 x 4 tiles (cluster split-K may multiply them); every CTA
    writes entry <= dependency wait passed <= accumulator in registers <= tile stored, an SM id below the SM count, and the
    call that switches the facility off reports one recorded launch."""
    L = _lib.lib()
    M, N, K = 1024, 512, 512
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); Cc = torch.empty(M, N, device="cuda")
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    def run():
        _lib.check(L.fb200_gemm(0, 1, M, N, K, Ct.c_void_p(A.data_ptr()), K, Ct.c_void_p(B.data_ptr()), K, Ct.c_void_p(Cc.data_ptr()), N,
                                None, 0, 0, Ct.c_void_p(ws.data_ptr()), ws.numel(), None), "fb200_gemm")
    run(); torch.cuda.synchronize()
    buf = torch.zeros(2 * 8192, dtype=torch.int64, device="cuda")
    L.fb200_debug_tc_timeline(Ct.c_void_p(buf.data_ptr()), 2)
    try:
        run(); torch.cuda.synchronize()
    finally:
        n = L.fb200_debug_tc_timeline(None, 0)
    assert n == 1
    a = buf.cpu().numpy().reshape(2, 1024, 8)[0]
    a = a[a[:, 0] > 0]
    assert len(a) >= 32 and len(a) % 32 == 0
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert (a[:, 3] >= 0).all() and (a[:, 3] < sms).all()
    assert (a[:, 0] <= a[:, 1]).all() and (a[:, 1] <= a[:, 4]).all() and (a[:, 4] <= a[:, 2]).all()
    assert (a[:, 7] <= a[:, 2]).all() and (a[:, 6] <= a[:, 5]).all()
    assert (a[:, 2] - a[:, 0]).max() < 1e6          # a tile lives microseconds, not milliseconds (globaltimer is in ns)
    assert parity.rel_err(Cc.cpu().numpy(), A.double().cpu().numpy() @ B.double().cpu().numpy().T) < 1e-5
    run(); torch.cuda.synchronize()                   # off again: nothing recorded
    assert L.fb200_debug_tc_timeline(None, 0) == 0
