"""Same-SM hand-over inside ONE tcgen05 GEMM launch with several waves of CTAs (fb200_debug_tc_timeline): the time from one CTA's
"tile stored" stamp to the entry of the next CTA on that SM.  usage: tc_handover.py engine layout M N K"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
from fusion_b200 import _lib
engine, layout, M, N, K = [int(v) for v in sys.argv[1:6]]
L = _lib.lib()
a_shape = (K, M) if layout == 2 else (M, K); b_shape = (N, K) if layout == 0 else (K, N)
A = torch.randn(*a_shape, device="cuda"); B = torch.randn(*b_shape, device="cuda"); Cc = torch.empty(M, N, device="cuda")
wsz = C.c_size_t(0); L.fb200_gemm_workspace_bytes(layout, engine, M, N, K, C.byref(wsz)); ws = torch.empty(max(wsz.value, 256), dtype=torch.uint8, device="cuda")
def run():
    return L.fb200_gemm(layout, engine, M, N, K, A.data_ptr(), a_shape[1], B.data_ptr(), b_shape[1], Cc.data_ptr(), N, None, 0, 0, ws.data_ptr(), ws.numel(), None)
for _ in range(3): run()
torch.cuda.synchronize()
buf = torch.zeros(2 * 8192, dtype=torch.int64, device="cuda")
L.fb200_debug_tc_timeline(buf.data_ptr(), 2); run(); torch.cuda.synchronize(); n = L.fb200_debug_tc_timeline(None, 0)
a = buf.cpu().numpy().reshape(2, 1024, 8)[0]; a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
ev = sorted((int(r[3]), (r[0] - t0) / 1e3, (r[1] - t0) / 1e3, (r[2] - t0) / 1e3) for r in a)
g = np.array([ev[i + 1][1] - ev[i][3] for i in range(len(ev) - 1) if ev[i][0] == ev[i + 1][0]] or [float("nan")])     # (one wave: no pairs)
life = np.array([e[3] - e[1] for e in ev])
m = lambda i, j: np.median(a[:, i] - a[:, j]) / 1e3
print(f"  relative to 'tile stored' (us, median): MMA warp left k-loop {m(7,2):+.2f}; accumulator in registers {m(4,2):+.2f}; dealloc issued {m(6,2):+.2f}; tensor memory released {m(5,2):+.2f}; entry {m(0,2):+.2f}; wait passed {m(1,2):+.2f}")
print(f"engine {engine} layout {layout} {M}x{N}x{K}: {len(a)} CTAs on {len(set(e[0] for e in ev))} SMs, CTA life median {np.median(life):.2f} us; "
      f"same-SM hand-over p10 {np.percentile(g,10):.2f} median {np.median(g):.2f} p90 {np.percentile(g,90):.2f} us ({len(g)} pairs); last stored {max(e[3] for e in ev):.1f} us")
