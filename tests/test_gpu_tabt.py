"""GPU: the fused TabTransformer (csrc/tabt.cu + fb200_linear_*, through the C ABI) against the float64 oracle that is pinned to
the reference class (tests/test_oracle_tabt.py): committed golden cases (the reference's own outputs), fresh inputs at the
reference's dimensions (82 columns, d = 32, 4 heads, ff = 128, 2 layers), eval and train with injected masks, plus
properties at B = 4096: determinism (bit-identical reruns) and Philox dropout statistics.  Tolerance: 1e-5 relative
(max-norm per tensor), the north star's fp32 bar."""
import os

import numpy as np
import pytest
import torch

import fusion_b200 as fb
from oracle import tabt_oracle as to
from tests import parity
from tests.golden import make_golden_tabt as G

pytestmark = pytest.mark.gpu
TOL = 1e-5
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tabt.npz"))


def build(cards, ncont, D, H, L, F, O, params, train, masks, dev="cuda"):
    m = fb.TabTransformer(cards, ncont, embed_dim=D, num_heads=H, num_transformer_layers=L, hidden_dim=F, output_dim=O, dropout=0.3)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()})
    m = m.to(dev).train(train)
    if masks is not None:
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        m._test_masks = {"enc": (t(masks["attn"]), t(masks["res1"]), t(masks["ff"]), t(masks["res2"])), "fc": t(masks["fc"])}
    return m


def run(m, x_cat, x_num, dout, dev="cuda"):
    xc = torch.from_numpy(x_cat).to(dev)
    xn = torch.from_numpy(x_num).float().to(dev).requires_grad_(x_num.shape[1] > 0)
    out = m(xc, xn)
    out.backward(torch.from_numpy(dout).float().to(dev))
    torch.cuda.synchronize()
    res = {"out": out.detach().cpu().numpy()}
    for k, p in m.named_parameters():
        res["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy()
    if x_num.shape[1] > 0:
        res["d_num"] = xn.grad.cpu().numpy()
    return res


def compare(got, ref, tol=TOL):
    worst = ("", 0.0)
    for k, v in got.items():
        r = ref[k]
        if np.abs(r).max() == 0:
            assert np.all(v == 0), k
            continue
        e = parity.rel_err(v, r)
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] < tol, worst
    return worst


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_matches_reference_golden(name):
    cards, ncont, D, H, L, F, O, B, train = G.CASES[name]
    params, x_cat, x_num, dout, masks = G.case_inputs(name)
    m = build(cards, ncont, D, H, L, F, O, params, train, masks)
    got = run(m, x_cat, x_num, dout)
    ref = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    compare(got, ref)


@pytest.mark.parametrize("B,train", [(7, False), (33, True), (200, True)])
def test_reference_dimensions_against_oracle(B, train):
    cards, ncont, D, H, L, F, O = [10] * 82, 4, 32, 4, 2, 128, 85            # loadImageModelClassifier.py:190-198
    rng = np.random.default_rng(B)
    params = to.gen_params(to.param_shapes(cards, ncont, D, F, L, O), 3)
    params = {k: v.astype(np.float32).astype(np.float64) for k, v in params.items()}
    x_cat = rng.integers(0, 10, size=(B, 82)).astype(np.int64)
    x_num = rng.standard_normal((B, ncont)).astype(np.float32).astype(np.float64)
    dout = rng.standard_normal((B, O)).astype(np.float32).astype(np.float64)
    masks = to.gen_masks(rng, L, B, 82, D, F, H, 0.3) if train else None
    for _ in range(20):
        # a ReLU pre-activation within fp32 rounding of zero (|.| < 2e-5 here) has no well-defined gradient in fp32: redraw
        # those samples' categories (the reference's own fp32 run would flip the same coins)
        r = to.forward_backward(params, x_cat, x_num, H, 0.3, masks, dout)
        tied = np.nonzero(r["sample_margin"] < 2e-5)[0]
        if tied.size == 0:
            break
        x_cat[tied] = rng.integers(0, 10, size=(tied.size, 82))
    assert tied.size == 0
    ref = {"out": r["out"], "d_num": r["d_num"], **{"grad/" + k: g for k, g in r["grads"].items()}}
    m = build(cards, ncont, D, H, L, F, O, params, train, masks)
    got = run(m, x_cat, x_num, dout)
    compare(got, ref)


def test_full_batch_is_deterministic_and_philox_dropout_has_the_right_rate():
    cards = [10] * 82
    torch.manual_seed(0)
    m = fb.TabTransformer(cards, 4, output_dim=85).cuda().train()
    B = 4096
    xc = torch.randint(0, 10, (B, 82), device="cuda")
    xn = torch.randn(B, 4, device="cuda")
    def step(offset):
        m.zero_grad(set_to_none=True)
        m._rng_calls = offset
        f = m.encode(xc, xn)
        f.square().sum().backward()
        # the encoder kernels' outputs: features, layer and embedding gradients (the numeric projection's weight gradient comes
        # from the FFMA split-K GEMM, which accumulates with atomics)
        return f.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()
                                    if p.grad is not None and (k.startswith("transformer_encoder") or k.startswith("embeddings"))}
    f1, g1 = step(5)
    assert len(g1) == 2 * 12 + 82
    f2, g2 = step(5)
    assert torch.equal(f1, f2)
    for k in g1:
        assert torch.equal(g1[k], g2[k]), k                               # no atomics anywhere: bit-reproducible
    f3, _ = step(6)
    assert not torch.equal(f1, f3)                                        # a new offset draws new masks
    m.eval()
    with torch.no_grad():
        e1 = m.encode(xc, xn); e2 = m.encode(xc, xn)
    assert torch.equal(e1, e2) and torch.isfinite(e1).all()
    # LayerNorm output: every token row of the encoder part is normalised before the affine map (default gamma = 1, beta = 0)
    tok = e1[:, :82 * 32].reshape(B, 82, 32)
    assert tok.mean(-1).abs().max() < 1e-5 and (tok.var(-1, unbiased=False) - 1).abs().max() < 1e-3


def test_config4_end_to_end_tab_transformer_feeds_the_fused_head():
    """BASELINE configs[3]: tab-transformer metadata + GFCAM fusion.  model(img_feat, (x_cat, x_num)) -> weighted CE ->
    backward through the fused head AND the fused TabTransformer, against the two oracles chained by hand the way SURVEY 8c
    prescribes (txt_feat = text_encoder(x_cat, x_num), then the reference's own sub-modules from text_projector on)."""
    from oracle import head_oracle as ho
    from tests.golden import cases as C
    B, F, Cn = 48, 768, 2
    kw = dict(mechanism="gfcam", F=F, V=None, C=Cn, T=85, text_model="tab-transformer")
    cfg = C.make_cfg(kw)
    hp = C.gen_params(cfg, 21, np.float32)
    cards = [10] * 82
    tp = {k: v.astype(np.float32) for k, v in to.gen_params(to.param_shapes(cards, 4, 32, 128, 2, 85), 22).items()}
    model = fb.MultimodalModel(Cn, cfg.H, "cuda", f"identity:{F}", "tab-transformer", common_dim=cfg.D, text_encoder_dim_output=85,
                               attention_mecanism="gfcam")
    sd = {k: torch.from_numpy(v) for k, v in hp.items()}
    sd.update({"text_encoder." + k: torch.from_numpy(v) for k, v in tp.items()})
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and not [k for k in res.missing_keys if not k.startswith("image_encoder.")], res
    model = model.cuda().eval()
    rng = np.random.default_rng(9)
    x = rng.standard_normal((B, F)).astype(np.float32)
    x_cat = rng.integers(0, 10, size=(B, 82)).astype(np.int64)
    x_num = rng.standard_normal((B, 4)).astype(np.float32)
    labels = rng.integers(0, Cn, size=B).astype(np.int64)
    cw = np.array([0.7, 1.6], np.float32)
    # oracles, float64, chained by hand; samples sitting on a ReLU tie of the encoder are redrawn
    tp64 = {k: v.astype(np.float64) for k, v in tp.items()}
    hp64 = {k: v.astype(np.float64) for k, v in hp.items()}
    for _ in range(20):
        enc = to.forward_backward(tp64, x_cat, x_num.astype(np.float64), 4, 0.3, None, None)
        tied = np.nonzero(enc["sample_margin"] < 2e-5)[0]
        if tied.size == 0:
            break
        x_cat[tied] = rng.integers(0, 10, size=(tied.size, 82))
    for _ in range(20):
        head = ho.head_forward_backward(cfg, hp64, x.astype(np.float64), enc["out"], labels, cw.astype(np.float64), None, need_input_grad=True)
        if head["relu_margin"] > 2e-5:
            break
        x = rng.standard_normal((B, F)).astype(np.float32)
    assert head["relu_margin"] > 2e-5
    encb = to.forward_backward(tp64, x_cat, x_num.astype(np.float64), 4, 0.3, None, head["d_text_in"])
    # CUDA: the drop-in call shape of the training loops
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    logits = model(xt, (torch.from_numpy(x_cat).cuda(), torch.from_numpy(x_num).cuda()))
    loss = fb.FusedCrossEntropyLoss(weight=torch.from_numpy(cw).cuda())(logits, torch.from_numpy(labels).cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert parity.rel_err(logits.detach().cpu().numpy(), head["logits"]) < TOL
    assert abs(float(loss) - head["loss"]) / abs(head["loss"]) < TOL
    assert np.array_equal(logits.argmax(1).cpu().numpy(), head["logits"].argmax(1))
    worst = ("", 0.0)
    named = dict(model.named_parameters())
    for k, g in head["grads"].items():
        if g is None or np.abs(g).max() == 0:
            continue
        e = parity.rel_err(named[k].grad.cpu().numpy(), g)
        worst = max(worst, (k, e), key=lambda kv: kv[1])
    for k, g in encb["grads"].items():
        if np.abs(g).max() == 0:
            continue
        e = parity.rel_err(named["text_encoder." + k].grad.cpu().numpy(), g)
        worst = max(worst, ("text_encoder." + k, e), key=lambda kv: kv[1])
    e = parity.rel_err(xt.grad.cpu().numpy(), head["d_img_feat"])
    worst = max(worst, ("d_img_feat", e), key=lambda kv: kv[1])
    assert worst[1] < 2e-5, worst          # two fp32 stages chained: the head's 1e-5 on top of the encoder's
