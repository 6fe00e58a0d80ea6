"""debug: fb200_gemm (tcgen05 3xTF32) over small shapes / epilogue flags against float64"""
import ctypes as Ct, itertools, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
from fusion_b200 import _lib
L = _lib.lib()
vp = lambda t: Ct.c_void_p(t.data_ptr())
bad = 0
from collections import Counter
cnt = Counter(); tot = Counter()
for engine in (1, 2):
  for layout, M, N, K, use_bias, relu, acc in itertools.product((0, 1), (32, 40, 128, 256), (256, 512), (128, 256, 512, 2048), (0, 1), (0, 1), (0, 1)):
    rng = np.random.default_rng(M + N + K + layout)
    a = rng.standard_normal((M, K)).astype(np.float32); b = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    A = torch.from_numpy(a).cuda(); Bm = torch.from_numpy(np.ascontiguousarray(b.T) if layout == 0 else b).cuda()
    Cc = torch.full((M, N), 0.5, device="cuda")
    wsz = Ct.c_size_t(0); _lib.check(L.fb200_gemm_workspace_bytes(layout, engine, M, N, K, Ct.byref(wsz)))
    ws = torch.empty(max(wsz.value, 256), dtype=torch.uint8, device="cuda")
    bt = torch.from_numpy(bias).cuda()
    _lib.check(L.fb200_gemm(layout, engine, M, N, K, vp(A), A.shape[1], vp(Bm), Bm.shape[1], vp(Cc), N, vp(bt) if use_bias else None, relu, acc, vp(ws), ws.numel(), None))
    torch.cuda.synchronize()
    if engine == 2:
        q = lambda x: torch.from_numpy(x).bfloat16().double().numpy(); ref = q(a) @ q(b)
    else:
        ref = a.astype(np.float64) @ b.astype(np.float64)
    if use_bias: ref = ref + bias
    if relu: ref = np.maximum(ref, 0)
    if acc: ref = ref + 0.5
    e = np.abs(Cc.cpu().numpy() - ref).max() / np.abs(ref).max()
    tot[(engine, layout, M, N, K)] += 1
    if not (e < (3e-6 if engine == 1 else 1e-5)):
        bad += 1; cnt[(engine, layout, M, N, K)] += 1
        if bad < 0: print("BAD engine", engine, "layout", layout, "M N K", M, N, K, "bias relu acc", use_bias, relu, acc, f"{e:.3e}")
print("bad cases:", bad)
for k in sorted(tot): print(k, "bad" if cnt[k] else "ok", cnt[k], "/", tot[k])
