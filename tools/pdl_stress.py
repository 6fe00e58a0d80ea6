"""PDL stress: the same forward+backward repeated many times must give the same gradients (up to atomic-order noise);
fused Adam on identical inputs must be bit-identical with PDL on and off."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
import fusion_b200 as fb
from fusion_b200 import _lib
from tests.golden import cases as C
from tests.gpu_util import build_model, case_inputs
L = _lib.lib()

def grads_once(model, x, tin, y, cw, crit, fused):
    model.zero_grad(set_to_none=True)
    if fused:
        loss, logits = model.forward_loss(x, tin, y, cw)
    else:
        logits = model(x, tin); loss = crit(logits, y); loss.backward()
    return float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

def stress(mech, B, dims, reps, fused):
    case = dict(cfg=dict(dims, mechanism=mech), B=B, seed=5, train=False, full_grads=False)
    cfg, model = build_model(case, "fp32"); x, tin, y, cw, _ = case_inputs(cfg, case); model.eval()
    crit = fb.FusedCrossEntropyLoss(weight=cw)
    out = {}
    for pdl in (0, 1):
        L.fb200_debug_set_pdl(pdl)
        l0, g0 = grads_once(model, x, tin, y, cw, crit, fused)
        worst = {}; lw = 0.0
        for _ in range(reps):
            l, g = grads_once(model, x, tin, y, cw, crit, fused)
            lw = max(lw, abs(l - l0) / abs(l0))
            for k in g0:
                d = float((g[k] - g0[k]).abs().max() / g0[k].abs().max().clamp_min(1e-30))
                worst[k] = max(worst.get(k, 0.0), d)
        torch.cuda.synchronize()
        bad = {k: v for k, v in worst.items() if v > 1e-5}
        out[pdl] = g0
        print(f"{mech[:20]:20s} B={B:4d} fused={int(fused)} pdl={pdl}: loss dev {lw:.2e}  max grad dev {max(worst.values()):.2e}  params over 1e-5: {bad}")
    dev = max(float((out[1][k] - out[0][k]).abs().max() / out[0][k].abs().max().clamp_min(1e-30)) for k in out[0])
    print(f"    pdl1 vs pdl0 first-run grads: {dev:.2e}")

def adam_stress(reps):
    torch.manual_seed(0)
    shapes = [(512, 2048), (512,), (6, 256), (3,), (1536, 512), (7, 13), (256, 85), (256,)]
    res = {}
    for pdl in (0, 1):
        L.fb200_debug_set_pdl(pdl)
        torch.manual_seed(1)
        ps = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
        opt = fb.FusedAdam(ps, lr=5e-3, weight_decay=1e-2)
        for it in range(reps):
            for p in ps: p.grad = torch.randn_like(p)     # torch kernel -> our adam kernel -> torch kernel ...
            opt.step()
        torch.cuda.synchronize()
        res[pdl] = [p.detach().clone() for p in ps]
    print("adam pdl1 == pdl0 bitwise:", all(torch.equal(a, b) for a, b in zip(res[0], res[1])))

if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    for fused in (False, True):
        stress("crossattention", 32, C.SMALL_DIMS, reps, fused)
        stress("att-intramodal+residual+cross-attention-metadados", 32, C.SMALL_DIMS, reps, fused)
        stress("crossattention", 256, dict(F=512, V=85, C=6), reps // 2, fused)
    adam_stress(100)
    L.fb200_debug_set_pdl(1)
