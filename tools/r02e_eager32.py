"""Eager forward_loss at B = 32 (the persistent step kernel route), ms per step; run under different FB200_* switches."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
import fusion_b200 as fb
dev = torch.device("cuda", 0)
torch.manual_seed(1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = fb.MultimodalModel(6, 8, dev, "identity:2048", "one-hot-encoder", vocab_size=85, text_encoder_dim_output=512, attention_mecanism="crossattention", compute_dtype="fp32").to(dev)
m.train()
x = torch.randn(B, 2048, device=dev); t = torch.randn(B, 85, device=dev); y = torch.randint(0, 6, (B,), device=dev); cw = torch.ones(6, device=dev)
for _ in range(20): m.forward_loss(x, t, y, cw)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(200): m.forward_loss(x, t, y, cw)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} eager forward_loss: {e0.elapsed_time(e1) / 200:.4f} ms per step (device), {(time.perf_counter() - t0) / 200 * 1e3:.4f} ms (host)")
