"""Pipeline trace of one tcgen05 GEMM CTA: prints per-k-block latencies (cycles)."""
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
nkb = (K + (31 if engine == 1 else 63)) // (32 if engine == 1 else 64)
tr = torch.zeros(8 * nkb + 64, dtype=torch.int64, device="cuda")
L.fb200_debug_tc_trace.argtypes = [C.c_void_p]
def run():
    return L.fb200_gemm(layout, engine, M, N, K, A.data_ptr(), a_shape[1], B.data_ptr(), b_shape[1], Cc.data_ptr(), N, None, 0, 0, ws.data_ptr(), ws.numel(), None)
for _ in range(3): run()
torch.cuda.synchronize()
L.fb200_debug_tc_trace(tr.data_ptr()); run(); torch.cuda.synchronize(); L.fb200_debug_tc_trace(None)
t = tr.cpu().numpy()[: 8 * nkb].reshape(nkb, 8)
t0 = t[0, 3]
S = 3 if engine == 1 else 6
print("kb   empty_seen  tma_issued  full/split_seen  mma_issued  split_done   (cycles since first)")
for i in range(min(nkb, 20)):
    print(f"{i:3d} {t[i,3]-t0:10d} {t[i,0]-t0:10d} {t[i,1]-t0:14d} {t[i,2]-t0:12d} {t[i,4]-t0 if t[i,4] else 0:12d}")
print("MMA warp gap (median): issued(i)->loop top(i+1)", np.median(t[5:, 7] - t[4:-1, 2]), " loop top->barrier passed", np.median(t[5:, 6] - t[5:, 7]), " fence", np.median(t[5:, 1] - t[5:, 6]))
print("MMA warp per k-block (median): split_seen->all issued+committed", np.median(t[4:, 2] - t[4:, 1]), " ->next split_seen", np.median(t[5:, 1] - t[4:-1, 2]))
print("entry->setup done", t[1,5]-t[0,5], " setup->first tma issued", t[0,0]-t[1,5], " first tma->first mma issued", t[0,2]-t[0,0], " last mma issued->acc in regs", t[2,5]-t[nkb-1,2], " acc in regs->parked in smem", t[4,5]-t[2,5], " parked->stored", t[3,5]-t[4,5], " total entry->stored", t[3,5]-t[0,5])
if t[5, 5]:      # cluster split-K kernel: own partial parked -> cluster barrier passed -> partial sums gathered (arrive) -> stored + cluster wait
    print("cluster epilogue: own tile parked->barrier passed", t[4,5]-t[5,5], " ->sums gathered", t[6,5]-t[4,5], " ->stored + peers done reading", t[3,5]-t[6,5])
print("median cycles between successive k-blocks at the MMA thread:", np.median(np.diff(t[:, 1])))
print("median tma_issued -> full/split seen:", np.median(t[:, 1] - t[:, 0]), " mma_issued(i) -> empty_seen(i+S):", np.median(t[S:, 3] - t[:-S, 2]))
