"""Grid-wide timeline of the tcgen05 GEMM launches of one train step (fb200_debug_tc_timeline): for every launch the entry, the
dependency-wait-passed and the tile-stored times of all its CTAs (globaltimer, ns), printed relative to the first entry.
usage: tc_timeline.py [workload] [batch] [dtype]"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
import fusion_b200 as fb
from fusion_b200 import _lib
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mech, F, V, Cn, T, tm, dtype = bench.WORKLOADS[wl]
if len(sys.argv) > 3: dtype = sys.argv[3]
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = fb.MultimodalModel(Cn, 8, dev, f"identity:{F}", "one-hot-encoder" if tm == 0 else "tab-transformer", vocab_size=V if V else 91,
                           text_encoder_dim_output=T, attention_mecanism=mech, compute_dtype=dtype).to(dev)
model.train()
x = torch.randn(B, F, device=dev); t = torch.randn(B, V if tm == 0 else T, device=dev); y = torch.randint(0, Cn, (B,), device=dev)
cw = torch.ones(Cn, device=dev)
L = _lib.lib()
for _ in range(3): model.forward_loss(x, t, y, cw)
torch.cuda.synchronize()
NL = 96
buf = torch.zeros(NL * 8192, dtype=torch.int64, device=dev)
L.fb200_debug_tc_timeline(buf.data_ptr(), NL)
model.forward_loss(x, t, y, cw)
torch.cuda.synchronize()
n = L.fb200_debug_tc_timeline(None, 0)
a = buf.cpu().numpy().reshape(NL, 1024, 8)
t0 = min(a[l][a[l][:, 0] > 0][:, 0].min() for l in range(n))
print(f"# {wl} B={B} {dtype}: {n} tcgen05 GEMM launches (host launch order; two lanes run concurrently); us since the first CTA entry")
print("# l  ctas | entry min..max | wait passed min..max | stored min..max | CTA life median (entry->stored) | wait median (entry->passed) | SMs")
rows = []
for l in range(n):
    c = a[l][a[l][:, 0] > 0]
    e, w, s = (c[:, 0] - t0) / 1e3, (c[:, 1] - t0) / 1e3, (c[:, 2] - t0) / 1e3
    rows.append((e, w, s))
    print(f"{l:3d} {len(c):5d} | {e.min():8.2f} {e.max():8.2f} | {w.min():8.2f} {w.max():8.2f} | {s.min():8.2f} {s.max():8.2f} | {np.median(s - e):6.2f} | {np.median(w - e):6.2f} | {len(set(c[:, 3]))}")
tot = max(r[2].max() for r in rows)
print(f"# last tile stored at {tot:.1f} us")

# per-SM hand-over: for every SM, the time from one CTA's "tile stored" stamp to the entry of the next CTA (of any recorded launch)
# on the same SM, when that next entry comes within 20 us
ev = []
for l in range(n):
    c = a[l][a[l][:, 0] > 0]
    for r in c: ev.append((int(r[3]), (r[0] - t0) / 1e3, (r[2] - t0) / 1e3, l))
ev.sort()
gaps = []
for i in range(len(ev) - 1):
    if ev[i][0] == ev[i + 1][0]:
        g = ev[i + 1][1] - ev[i][2]
        if 0 <= g < 20: gaps.append((g, ev[i][3], ev[i + 1][3]))
g = np.array([x[0] for x in gaps])
print(f"# same-SM hand-over (tile stored -> next CTA entry), {len(g)} pairs: p10 {np.percentile(g,10):.2f}  median {np.median(g):.2f}  p90 {np.percentile(g,90):.2f}  max {g.max():.2f} us")
same = np.array([x[0] for x in gaps if x[2] == x[1] + 1]); other = np.array([x[0] for x in gaps if x[2] != x[1] + 1])
if len(same): print(f"#   next CTA from the following launch: median {np.median(same):.2f} ({len(same)} pairs);  from another launch: median {np.median(other) if len(other) else float('nan'):.2f} ({len(other)} pairs)")
for l in range(min(n, 12)):
    c = a[l][a[l][:, 0] > 0]
    e = np.sort((c[:, 0] - t0) / 1e3); s = np.sort((c[:, 2] - t0) / 1e3)
    print(f"#   launch {l}: entry percentiles 0/25/50/75/100 = {e[0]:.1f} {e[len(e)//4]:.1f} {e[len(e)//2]:.1f} {e[3*len(e)//4]:.1f} {e[-1]:.1f}   stored = {s[0]:.1f} {s[len(s)//4]:.1f} {s[len(s)//2]:.1f} {s[3*len(s)//4]:.1f} {s[-1]:.1f}")
np.save(os.path.join(ROOT, "gpurun_out", f"r02e_timeline_{wl}_B{B}.npy"), a[:n])
