"""One forward + backward of the fused TabTransformer encoder at the reference dimensions (the command ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
import fusion_b200 as fb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = fb.TabTransformer([10] * 82, 4, output_dim=85).cuda().train()
xc = torch.randint(0, 10, (B, 82), device="cuda"); xn = torch.randn(B, 4, device="cuda")
for _ in range(2):
    m.zero_grad(set_to_none=True)
    m(xc, xn).square().sum().backward()
torch.cuda.synchronize()
print("ok")
