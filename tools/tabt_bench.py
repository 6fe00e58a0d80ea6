"""TabTransformer micro-benchmark on one GPU: the fused kernels (fusion_b200.TabTransformer) against the same module built
from stock torch.nn (the incumbent: what models/tab_transformer.py:6-60 runs on device='cuda'), forward + backward,
reference dimensions (82 columns x 10 categories, d = 32, 4 heads, ff = 128, 2 layers, 4 numeric columns, 85 outputs).

    python tools/tabt_bench.py [B ...]        # default 32 1024 4096
"""
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import fusion_b200 as fb  # noqa: E402


class StockTabTransformer(nn.Module):
    """torch.nn composition with the reference's structure (the incumbent on the same GPU)."""

    def __init__(self, cards, ncont, D=32, H=4, L=2, F=128, O=1, p=0.3):
        super().__init__()
        self.embeddings = nn.ModuleList([nn.Embedding(c, D) for c in cards])
        layer = nn.TransformerEncoderLayer(d_model=D, nhead=H, dim_feedforward=F, activation="relu", dropout=p, batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=L)
        self.numeric_projection = nn.Linear(ncont, D)
        self.fc = nn.Sequential(nn.Linear(len(cards) * D + D, F), nn.ReLU(), nn.Dropout(p), nn.Linear(F, O))

    def forward(self, xc, xn):
        tok = torch.stack([e(xc[:, i]) for i, e in enumerate(self.embeddings)], dim=1)
        return self.fc(torch.cat([self.transformer_encoder(tok).flatten(start_dim=1), self.numeric_projection(xn)], dim=1))


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
    batches = [int(a) for a in sys.argv[1:]] or [32, 1024, 4096]
    torch.backends.cuda.matmul.allow_tf32 = False
    cards = [10] * 82
    out = {}
    for B in batches:
        torch.manual_seed(0)
        fused = fb.TabTransformer(cards, 4, output_dim=85).cuda().train()
        stock = StockTabTransformer(cards, 4, O=85).cuda().train()
        xc = torch.randint(0, 10, (B, 82), device="cuda")
        xn = torch.randn(B, 4, device="cuda")
        g = torch.randn(B, 85, device="cuda")

        def step(m):
            def f():
                for p in m.parameters():
                    p.grad = None
                m(xc, xn).backward(g)
            return f

        def fwd(m):
            def f():
                with torch.no_grad():
                    m(xc, xn)
            return f

        def enc_only():
            for p in fused.parameters():
                p.grad = None
            fused.encode(xc, xn).backward(gf)
        gf = torch.randn(B, 82 * 32 + 32, device="cuda")
        r = {"fused_fwd_bwd_ms": time_ms(step(fused)), "stock_fwd_bwd_ms": time_ms(step(stock)),
             "fused_encoder_fwd_bwd_ms": time_ms(enc_only)}
        fused.eval(); stock.eval()
        r["fused_eval_fwd_ms"] = time_ms(fwd(fused)); r["stock_eval_fwd_ms"] = time_ms(fwd(stock))
        r["speedup_train"] = r["stock_fwd_bwd_ms"] / r["fused_fwd_bwd_ms"]
        r["speedup_eval"] = r["stock_eval_fwd_ms"] / r["fused_eval_fwd_ms"]
        # algorithmic FLOPs of the encoder stack per sample (forward; backward with recompute = 3x this + the recompute)
        T, D, H, F, L = 82, 32, 4, 128, 2
        fl = L * (2 * T * D * 3 * D + 4 * T * T * D + 2 * T * D * D + 4 * T * D * F)
        r["encoder_fwd_gflop"] = fl * B / 1e9
        out[B] = r
        print(B, json.dumps(r), flush=True)
    print(json.dumps({"tabt_bench": out}))


if __name__ == "__main__":
    main()
