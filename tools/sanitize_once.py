"""Small end-to-end exercise for compute-sanitizer (memcheck): the fused head on the tcgen05 path (B = 160, two lanes, PDL)
and on the FFMA path (B = 8), the fused Adam and the token attention with ragged tiles."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
import fusion_b200 as fb
torch.manual_seed(0)
for B in (160, 8):
    m = fb.MultimodalModel(6, 8, "cuda", "identity:512", "one-hot-encoder", vocab_size=85, attention_mecanism="gfcam").cuda()
    m.train()
    opt = fb.FusedAdam(m.parameters(), lr=1e-3)
    x = torch.randn(B, 512, device="cuda"); t = torch.randn(B, 85, device="cuda"); y = torch.randint(0, 6, (B,), device="cuda")
    loss, _ = m.forward_loss(x, t, y, torch.ones(6, device="cuda"))
    opt.step()
    out = m(x, t); fb.FusedCrossEntropyLoss()(out, y).backward()
att = fb.MultiheadAttention(64, 4).cuda()
q = torch.randn(19, 3, 64, device="cuda", requires_grad=True); kv = torch.randn(37, 3, 64, device="cuda", requires_grad=True)
o, _ = att(q, kv, kv); o.sum().backward()
torch.cuda.synchronize()
print("ok", float(loss))
