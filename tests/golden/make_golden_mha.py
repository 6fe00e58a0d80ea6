"""Golden vectors for multi-head attention on token sequences: outputs and gradients of
torch.nn.MultiheadAttention (float64, CPU) - the module the reference instantiates
(models/multimodalIntraInterModal.py:78-100, models/multimodalGated.py:118-206) and the library that owns its
arithmetic.  Run in the build container:  python tests/golden/make_golden_mha.py  -> tests/golden/mha.npz"""
import os

import numpy as np
import torch

CASES = {  # name: (Sq, Skv, B, D, H, self_attention)
    "s1": (1, 1, 4, 32, 4, False),
    "tiny": (5, 7, 3, 32, 4, False),
    "hd32": (33, 20, 2, 64, 2, False),
    "self": (19, 19, 2, 32, 2, True),
}


def run_torch(Sq, Sk, B, D, H, self_attn, seed):
    g = torch.Generator().manual_seed(seed)
    m = torch.nn.MultiheadAttention(D, H).double()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float64) * 0.2)
    q = torch.randn(Sq, B, D, generator=g, dtype=torch.float64, requires_grad=True)
    if self_attn:
        k = v = q
    else:
        k = torch.randn(Sk, B, D, generator=g, dtype=torch.float64, requires_grad=True)
        v = torch.randn(Sk, B, D, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(Sq, B, D, generator=g, dtype=torch.float64)
    out, _ = m(q, k, v)
    out.backward(dy)
    r = dict(q=q, k=k, v=v, dy=dy, out=out, in_w=m.in_proj_weight, in_b=m.in_proj_bias, out_w=m.out_proj.weight, out_b=m.out_proj.bias,
             dq=q.grad, d_in_w=m.in_proj_weight.grad, d_in_b=m.in_proj_bias.grad, d_out_w=m.out_proj.weight.grad, d_out_b=m.out_proj.bias.grad)
    if not self_attn:
        r.update(dk=k.grad, dv=v.grad)
    return {n: t.detach().numpy().copy() for n, t in r.items()}


if __name__ == "__main__":
    blob = {}
    for i, (name, (Sq, Sk, B, D, H, sa)) in enumerate(CASES.items()):
        for k, v in run_torch(Sq, Sk, B, D, H, sa, 100 + i).items():
            blob[f"{name}/{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mha.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")
