"""GPU probe for the tcgen05 GEMM: one (engine, layout, M, N, K) per process so that a protocol
bug (trap / timeout) cannot poison the other cases.  Usage: tc_probe.py engine layout M N K"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
from fusion_b200 import _lib  # noqa: E402


def main():
    engine, layout, M, N, K = [int(v) for v in sys.argv[1:6]]
    extra = sys.argv[6] if len(sys.argv) > 6 else ""
    L = _lib.lib()
    rng = np.random.default_rng(M * 7 + N * 3 + K + layout)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    A = torch.from_numpy(a if layout != 2 else np.ascontiguousarray(a.T)).cuda()
    Bm = torch.from_numpy(np.ascontiguousarray(b.T) if layout == 0 else b).cuda()
    Cc = torch.full((M, N), 0.25, device="cuda")
    wsz = C.c_size_t(0)
    _lib.check(L.fb200_gemm_workspace_bytes(layout, engine, M, N, K, C.byref(wsz)), "ws")
    ws = torch.empty(wsz.value, dtype=torch.uint8, device="cuda")
    use_bias = extra != "nobias"
    bt = torch.from_numpy(bias).cuda()
    rc = L.fb200_gemm(layout, engine, M, N, K, A.data_ptr(), A.shape[1], Bm.data_ptr(), Bm.shape[1], Cc.data_ptr(), N,
                      bt.data_ptr() if use_bias else None, 1 if use_bias else 0, 1 if use_bias else 0, ws.data_ptr(), ws.numel(), None)
    torch.cuda.synchronize()
    if rc != 0:
        print(f"RC {rc}"); sys.exit(2)
    if engine == 2:
        bf = lambda x: torch.from_numpy(x).bfloat16().double().numpy()
        ref = bf(a) @ bf(b)
    else:
        ref = a.astype(np.float64) @ b.astype(np.float64)
    if use_bias:
        ref = np.maximum(ref + bias, 0) + 0.25
    got = Cc.cpu().numpy().astype(np.float64)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    # timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        L.fb200_gemm(layout, engine, M, N, K, A.data_ptr(), A.shape[1], Bm.data_ptr(), Bm.shape[1], Cc.data_ptr(), N, None, 0, 0, ws.data_ptr(), ws.numel(), None)
    e0.record()
    for _ in range(10):
        L.fb200_gemm(layout, engine, M, N, K, A.data_ptr(), A.shape[1], Bm.data_ptr(), Bm.shape[1], Cc.data_ptr(), N, None, 0, 0, ws.data_ptr(), ws.numel(), None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"engine {engine} layout {layout} M {M} N {N} K {K} {extra}: rel err {err:.3e}  {ms * 1e3:.1f} us incl. operand conversion  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    tol = 3e-6 if engine == 1 else 1e-5
    sys.exit(0 if err < tol else 1)


if __name__ == "__main__":
    main()
