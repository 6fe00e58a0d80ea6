"""Phase timeline of the TabTransformer kernels: clock64 stamps of CTA 0 after every __syncthreads-separated phase of its first
sample (forward: layer 0; backward: the top layer = recompute + backward phases).   python tools/tabt_trace.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
import fusion_b200 as fb
from fusion_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
m = fb.TabTransformer([10] * 82, 4, output_dim=85).cuda().train()
xc = torch.randint(0, 10, (B, 82), device="cuda"); xn = torch.randn(B, 4, device="cuda")
g = torch.randn(B, 82 * 32 + 32, device="cuda")
for _ in range(3):
    m.zero_grad(set_to_none=True); m.encode(xc, xn).backward(g)
torch.cuda.synchronize()
tr = torch.zeros(32, dtype=torch.int64, device="cuda")
L = _lib.lib()
FWD = ["in-proj (QKV)", "attention rows", "out-proj", "residual+LN1", "ff linear1+relu+drop", "ff linear2", "residual+LN2"]
BWD = ["LN2 param grads", "LN2 bwd", "dW2 + db2", "dH (in place over H)", "dW1 + db1 + dX1", "LN1 param grads", "LN1 bwd", "dWo + dbo + dA", "attention dQ", "attention dK dV", "dWin + dbin + dX"]
L.fb200_debug_tabt_trace(tr.data_ptr())
with torch.no_grad():
    m.eval(); m.encode(xc, xn); m.train()
torch.cuda.synchronize()
t = tr.cpu().numpy().copy()
print(f"forward kernel, CTA 0, first sample, layer 0 (cycles; B = {B}, {os.environ.get('TABT_NOTE', '')})")
for i, n in enumerate(FWD):
    print(f"  {n:24s} {t[i + 1] - t[i]:8d}")
print(f"  {'layer total':24s} {t[7] - t[0]:8d}")
tr.zero_()
f = m.encode(xc, xn); torch.cuda.synchronize()
tr.zero_()
f.backward(g); torch.cuda.synchronize()
L.fb200_debug_tabt_trace(None)
t = tr.cpu().numpy().copy()
print("backward kernel, CTA 0, first sample, top layer (cycles)")
for i, n in enumerate(FWD):
    print(f"  recompute {n:24s} {t[i + 1] - t[i]:8d}")
for i, n in enumerate(BWD):
    print(f"  {n:34s} {t[8 + i] - t[7 + i]:8d}")
print(f"  {'layer total':34s} {t[18] - t[0]:8d}")
