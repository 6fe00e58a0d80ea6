"""Token attention micro-benchmark on one GPU (197 image tokens attending to 85 metadata tokens, D = 512, 8 heads - the
sequence shape of models/multimodalGated.py:118-206): fusion_b200.MultiheadAttention forward + backward against
torch.nn.MultiheadAttention on the same device (fp32, TF32 off), core kernels timed inside by CUDA events.

    python tools/attn_bench.py [B]            # FB200_ATTN_TC=0 selects the FFMA core for A/B
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import fusion_b200 as fb  # noqa: E402


def time_ms(fn, warm=5, it=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    Sq, Sk, D, H = 197, 85, 512, 8
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    mine = fb.MultiheadAttention(D, H).cuda()
    pooled = fb.MultiheadAttention(D, H, pool="mean").cuda()
    ref = torch.nn.MultiheadAttention(D, H).cuda()
    ref.load_state_dict(mine.state_dict())
    pooled.load_state_dict(mine.state_dict())
    q = torch.randn(Sq, B, D, device="cuda", requires_grad=True)
    kv = torch.randn(Sk, B, D, device="cuda", requires_grad=True)
    g = torch.randn(Sq, B, D, device="cuda")
    gp = torch.randn(B, D, device="cuda")

    def step(m, grad):
        def f():
            q.grad = kv.grad = None
            for p in m.parameters():
                p.grad = None
            out, _ = m(q, kv, kv)
            out.backward(grad)
        return f

    def fwd(m):
        def f():
            with torch.no_grad():
                m(q, kv, kv)
        return f

    a, _ = mine(q, kv, kv); b, _ = ref(q, kv, kv)
    err = float((a - b).abs().max() / b.abs().max())
    core_flops = 4.0 * Sq * Sk * D * B                      # QK^T and PV, forward
    r = {"B": B, "Sq": Sq, "Sk": Sk, "D": D, "H": H, "tc_core": os.environ.get("FB200_ATTN_TC", "1") != "0",
         "fused_fwd_bwd_ms": time_ms(step(mine, g)), "torch_fwd_bwd_ms": time_ms(step(ref, g)),
         "fused_pooled_fwd_bwd_ms": time_ms(step(pooled, gp)),
         "fused_fwd_ms": time_ms(fwd(mine)), "torch_fwd_ms": time_ms(fwd(ref)), "max_rel_diff_vs_torch": err,
         "core_fwd_gflop": core_flops / 1e9}
    r["speedup_fwd_bwd"] = r["torch_fwd_bwd_ms"] / r["fused_fwd_bwd_ms"]
    print(json.dumps(r))


if __name__ == "__main__":
    main()
