"""One eager train step (forward + CE + backward, no CUDA graph) of the workloads whose HBM-bound row kernels the north star
names - LayerNorm+ReLU+dropout, gates, gated residual, MetaBlock, cross entropy, classifier head - at B = 4096 fp32, for
`ncu --set full -k regex:...` (profiles/README.md has the command).  Usage: python tools/profile_row_kernels.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200")]
import torch
import fusion_b200 as fb
from bench import WORKLOADS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
for wl in ("cfg2", "cfg3a", "cfg3b", "cfg5"):
    mech, F, V, Cn, T, tm, _ = WORKLOADS[wl]
    torch.manual_seed(1234)
    model = fb.MultimodalModel(Cn, 8, dev, f"identity:{F}", "one-hot-encoder", vocab_size=V, text_encoder_dim_output=T,
                               attention_mecanism=mech, compute_dtype="fp32").to(dev).train()
    x = torch.randn(B, F, device=dev); t = torch.randn(B, V, device=dev); y = torch.randint(0, Cn, (B,), device=dev)
    cw = torch.ones(Cn, device=dev)
    for _ in range(int(os.environ.get("FB200_PROFILE_STEPS", "2"))):
        loss, _ = model.forward_loss(x, t, y, cw)
    torch.cuda.synchronize()
    print(wl, mech, "B", B, "loss", float(loss))
