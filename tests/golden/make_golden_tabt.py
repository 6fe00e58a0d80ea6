"""Golden vectors for the TabTransformer oracle, produced by the UNMODIFIED reference class
(/root/reference/src/scripts/benchmark/models/tab_transformer.py:6-60) in float64 on CPU.

    python tests/golden/make_golden_tabt.py        # rewrites tests/golden/tabt.npz (needs /root/reference)

Eval cases run the class as is.  Train cases need known dropout masks: torch.nn.functional.dropout (nn.Dropout modules of the
layer and of the fc MLP) and torch.nn.functional.scaled_dot_product_attention (the attention-probability dropout inside
nn.MultiheadAttention's need_weights=False path) are swapped, for the duration of the call, for versions that take their
keep-masks from a queue in call order; everything else is the reference's own code.
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/src/scripts/benchmark/models/tab_transformer.py"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name: (cardinalities, num_continuous, embed_dim, heads, layers, hidden, output_dim, B, train)
CASES = {
    "eval_small": ([5, 3, 7, 4, 6, 2, 9], 3, 16, 4, 2, 32, 11, 5, False),
    "train_small": ([5, 3, 7, 4, 6, 2, 9], 3, 16, 4, 2, 32, 11, 5, True),
    "eval_ref_dims": ([10] * 12, 4, 32, 4, 2, 128, 85, 3, False),       # the reference's layer dims (loadImageModelClassifier.py:190-198), 12 columns
    "train_one_layer": ([4, 4, 4], 0, 8, 2, 1, 16, 3, 4, True),       # no numeric columns: numeric_projection is None (:30)
}


def reference_available():
    return os.path.isfile(REF)


def _load_reference_class():
    spec = importlib.util.spec_from_file_location("ref_tab_transformer", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TabTransformer


def case_inputs(name, seed=0):
    from oracle import tabt_oracle as to
    cards, ncont, D, H, L, F, O, B, train = CASES[name]
    rng = np.random.default_rng(1000 + seed + len(name))
    params = to.gen_params(to.param_shapes(cards, ncont, D, F, L, O), seed + 17)
    x_cat = np.stack([rng.integers(0, c, size=B) for c in cards], axis=1).astype(np.int64)
    x_num = rng.standard_normal((B, ncont))
    dout = rng.standard_normal((B, O))
    masks = to.gen_masks(rng, L, B, len(cards), D, F, H, 0.3) if train else None
    return params, x_cat, x_num, dout, masks


def run_reference(name, seed=0):
    """-> dict of float64 arrays: out, grad/<param name>, d_num."""
    import torch
    import torch.nn.functional as Fn
    cards, ncont, D, H, L, F, O, B, train = CASES[name]
    params, x_cat, x_num, dout, masks = case_inputs(name, seed)
    Ref = _load_reference_class()
    torch.manual_seed(0)
    m = Ref(cards, ncont, embed_dim=D, num_heads=H, num_transformer_layers=L, hidden_dim=F, output_dim=O, dropout=0.3).double()
    sd = m.state_dict()
    assert list(sd.keys()) == list(params.keys()), (list(sd.keys()), list(params.keys()))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    m.train(train)
    xc, xn = torch.from_numpy(x_cat), torch.from_numpy(x_num).requires_grad_(ncont > 0)
    orig_dropout, orig_sdpa = Fn.dropout, Fn.scaled_dot_product_attention
    if train:
        queue = []
        for l in range(L):                   # call order inside one layer: attention probabilities, dropout1, ff dropout, dropout2
            queue += [("attn", masks["attn"][l]), ("drop", masks["res1"][l]), ("drop", masks["ff"][l]), ("drop", masks["res2"][l])]
        queue.append(("drop", masks["fc"]))

        def fake_dropout(x, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return x
            kind, mk = queue.pop(0)
            assert kind == "drop" and tuple(mk.shape) == tuple(x.shape), (kind, mk.shape, x.shape)
            return x * torch.from_numpy(mk.astype(np.float64)) / (1.0 - p)

        def fake_sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, **kw):
            assert attn_mask is None and not is_causal
            w = torch.softmax(q @ k.transpose(-2, -1) / np.sqrt(q.shape[-1]), dim=-1)
            if dropout_p > 0.0:
                kind, mk = queue.pop(0)
                assert kind == "attn" and tuple(mk.shape) == tuple(w.shape), (kind, mk.shape, w.shape)
                w = w * torch.from_numpy(mk.astype(np.float64)) / (1.0 - dropout_p)
            return w @ v

        Fn.dropout, Fn.scaled_dot_product_attention = fake_dropout, fake_sdpa
    try:
        out = m(xc, xn)
        out.backward(torch.from_numpy(dout))
    finally:
        Fn.dropout, Fn.scaled_dot_product_attention = orig_dropout, orig_sdpa
    if train:
        assert not queue, f"{len(queue)} masks were not consumed"
    res = {"out": out.detach().numpy().copy()}
    for k, p in m.named_parameters():
        res["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    if ncont > 0:
        res["d_num"] = xn.grad.numpy().copy()
    return res


def main():
    blob = {}
    for name in CASES:
        for k, v in run_reference(name).items():
            blob[f"{name}/{k}"] = v
    path = os.path.join(HERE, "tabt.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
