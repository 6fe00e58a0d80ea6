"""Per-stage timing of the persistent step kernel (csrc/mega.cuh): clock64 stamps of CTA 0 around every stage, next to
the stage table (FB200_MEGA_DUMP=1), and the cost of the grid barriers alone.  Run on a B200:
    FB200_MEGA_DUMP=1 python tools/mega_trace.py [workload] [batch]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200")]
import torch
import fusion_b200 as fb
from fusion_b200 import _lib
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
mech, F, V, Cn, T, tm, dtype = WORKLOADS[wl]
dev = torch.device("cuda", 0)
L = _lib.lib()
MHZ = 1965.0

# ---- barriers alone
ws = torch.zeros(64, dtype=torch.int32, device=dev)
for n in (1, 2, 11, 41):
    for _ in range(3):
        L.fb200_debug_mega_barriers(n, ws.data_ptr(), None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.fb200_debug_mega_barriers(n, ws.data_ptr(), None)
    e1.record(); torch.cuda.synchronize()
    print(f"empty step kernel, {n:2d} stages: {e0.elapsed_time(e1) / 20 * 1e3:8.2f} us per launch (memset + cooperative launch + {n - 1} grid barriers)")

torch.manual_seed(1234)
model = fb.MultimodalModel(Cn, 8, dev, f"identity:{F}", "one-hot-encoder" if tm == 0 else "tab-transformer", vocab_size=V if V else 91,
                           text_encoder_dim_output=T, attention_mecanism=mech, compute_dtype="fp32").to(dev).train()
x = torch.randn(B, F, device=dev); t = torch.randn(B, V if tm == 0 else T, device=dev); y = torch.randint(0, Cn, (B,), device=dev)
cw = torch.ones(Cn, device=dev)
for _ in range(3):
    model.forward_loss(x, t, y, cw)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
L.fb200_debug_mega_trace(buf.data_ptr())
model.forward_loss(x, t, y, cw)
torch.cuda.synchronize()
L.fb200_debug_mega_trace(None)
tr = buf.cpu().tolist()
t0 = tr[0]
print(f"{wl} B={B}: stage | own tasks done (us since entry) | barrier passed | stage total")
prev = t0
s = 0
while 2 + 2 * s < len(tr) and (tr[1 + 2 * s] or tr[2 + 2 * s]):
    a, b = tr[1 + 2 * s], tr[2 + 2 * s] or tr[1 + 2 * s]
    print(f"  {s:2d} | {(a - t0) / MHZ:8.2f} | {(b - t0) / MHZ:8.2f} | work {(a - prev) / MHZ:6.2f} + wait {(b - a) / MHZ:6.2f}")
    d = tr[256 + 8 * s: 256 + 8 * s + 8]
    if d[4]:
        print(f"       stage top {(d[3] - t0) / MHZ:8.2f} | gemm op begin +{(d[4] - d[3]) / MHZ:5.2f} | NT task entry +{(d[0] - d[4]) / MHZ:5.2f} | k loop {(d[1] - d[0]) / MHZ:5.2f} | reduce {(d[2] - d[1]) / MHZ:5.2f} | op end +{(d[5] - d[2]) / MHZ:5.2f}" if d[0] else
              f"       stage top {(d[3] - t0) / MHZ:8.2f} | gemm op begin +{(d[4] - d[3]) / MHZ:5.2f} | op total {(d[5] - d[4]) / MHZ:5.2f}")
    prev = b; s += 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    model.forward_loss(x, t, y, cw)
e1.record(); torch.cuda.synchronize()
print(f"eager forward_loss: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per step")
