"""debug: structure of the error of one cluster-split layout-1 fp32 GEMM"""
import ctypes as Ct, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
from fusion_b200 import _lib
L = _lib.lib()
vp = lambda t: Ct.c_void_p(t.data_ptr())
layout, engine = int(sys.argv[1]), 1
M, N, K = 128, 256, 512
rng = np.random.default_rng(1)
a = rng.standard_normal((M, K)).astype(np.float32); b = rng.standard_normal((K, N)).astype(np.float32)
A = torch.from_numpy(a).cuda(); Bm = torch.from_numpy(np.ascontiguousarray(b.T) if layout == 0 else b).cuda()
Cc = torch.full((M, N), 0.5, device="cuda")
ws = torch.empty(256, dtype=torch.uint8, device="cuda")
_lib.check(L.fb200_gemm(layout, engine, M, N, K, vp(A), A.shape[1], vp(Bm), Bm.shape[1], vp(Cc), N, None, 0, 0, vp(ws), ws.numel(), None))
torch.cuda.synchronize()
got = Cc.cpu().numpy().astype(np.float64)
ref = a.astype(np.float64) @ b.astype(np.float64)
err = np.abs(got - ref) / np.abs(ref).max()
print("layout", layout, "max err", err.max(), "nan", np.isnan(got).sum(), "inf", np.isinf(got).sum())
print("bad fraction by row%8:", [float((err[r::8] > 1e-5).mean().round(3)) for r in range(8)])
print("bad fraction by col block of 32:", [float((err[:, c:c + 32] > 1e-5).mean().round(3)) for c in range(0, N, 32)])
# which k-slices are present?  least squares of got against the per-slice partial products
S = 4
parts = [a[:, s * K // S:(s + 1) * K // S].astype(np.float64) @ b[s * K // S:(s + 1) * K // S].astype(np.float64) for s in range(S)]
for r in (0, 1, 2, 3):
    Xm = np.stack([p[r::4].ravel() for p in parts], 1)
    y = np.nan_to_num(got[r::4].ravel(), posinf=0, neginf=0)
    coef, *_ = np.linalg.lstsq(Xm, y, rcond=None)
    print("rows%4 ==", r, "coefficients of the 4 k-slice partials:", coef.round(3), "residual", float(np.abs(Xm @ coef - y).max().round(3)))
