"""One forward + backward of the fused token attention (197 image tokens x 85 metadata tokens, batch 32, D 512, 8 heads):
the command ncu captures for profiles/r01_attention_*."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
import fusion_b200 as fb
torch.manual_seed(0)
m = fb.MultiheadAttention(512, 8).cuda()
q = torch.randn(197, 32, 512, device="cuda", requires_grad=True); kv = torch.randn(85, 32, 512, device="cuda", requires_grad=True)
for _ in range(2):
    out, _ = m(q, kv, kv); out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
