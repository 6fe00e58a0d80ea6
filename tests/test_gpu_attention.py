"""GPU: the fused multi-head attention (csrc/attention.cuh through fb200_mha_forward / fb200_mha_backward) against the
float64 oracle (oracle/head_oracle.py: mha, pinned to torch.nn.MultiheadAttention by tests/test_oracle_mha.py), the
committed golden vectors, and torch's own module on the GPU.  fp32 tolerance: 1e-5 relative (max norm)."""
import os

import numpy as np
import pytest
import torch

import fusion_b200 as fb
from tests import parity
from tests.golden import make_golden_mha as G
from tests.test_oracle_mha import GOLD, oracle_mha

pytestmark = pytest.mark.gpu
TOL = 1e-5


def run_cuda(c, H, self_attn):
    dev = "cuda"
    D = c["q"].shape[-1]
    m = fb.MultiheadAttention(D, H).to(dev)
    with torch.no_grad():
        m.in_proj_weight.copy_(torch.from_numpy(c["in_w"])); m.in_proj_bias.copy_(torch.from_numpy(c["in_b"]))
        m.out_proj.weight.copy_(torch.from_numpy(c["out_w"])); m.out_proj.bias.copy_(torch.from_numpy(c["out_b"]))
    q = torch.from_numpy(c["q"]).float().to(dev).requires_grad_(True)
    if self_attn:
        k = v = q
    else:
        k = torch.from_numpy(c["k"]).float().to(dev).requires_grad_(True)
        v = torch.from_numpy(c["v"]).float().to(dev).requires_grad_(True)
    out, w = m(q, k, v)
    assert w is None
    out.backward(torch.from_numpy(c["dy"]).float().to(dev))
    torch.cuda.synchronize()
    r = dict(out=out, dq=q.grad, d_in_w=m.in_proj_weight.grad, d_in_b=m.in_proj_bias.grad, d_out_w=m.out_proj.weight.grad, d_out_b=m.out_proj.bias.grad)
    if not self_attn:
        r.update(dk=k.grad, dv=v.grad)
    return {n: t.detach().cpu().numpy() for n, t in r.items()}


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_golden(name):
    Sq, Sk, B, D, H, sa = G.CASES[name]
    c = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/")}
    got = run_cuda(c, H, sa)
    for k, v in got.items():
        if np.abs(c[k]).max() == 0:
            assert np.all(v == 0), k          # S = 1: W_q / W_k / query gradients are exact zeros
        else:
            assert parity.rel_err(v, c[k]) < TOL, (k, parity.rel_err(v, c[k]))


def _fresh(Sq, Sk, B, D, H, sa, seed):
    rng = np.random.default_rng(seed)
    c = dict(q=rng.standard_normal((Sq, B, D)), dy=rng.standard_normal((Sq, B, D)),
             in_w=rng.standard_normal((3 * D, D)) / np.sqrt(D), in_b=0.1 * rng.standard_normal(3 * D),
             out_w=rng.standard_normal((D, D)) / np.sqrt(D), out_b=0.1 * rng.standard_normal(D))
    c["k"] = c["q"] if sa else rng.standard_normal((Sk, B, D))
    c["v"] = c["q"] if sa else rng.standard_normal((Sk, B, D))
    return {k: v.astype(np.float32).astype(np.float64) for k, v in c.items()}     # fp32-representable inputs


@pytest.mark.parametrize("shape", [
    (50, 85, 2, 512, 8, False),      # hd = 64: image tokens attending to metadata tokens at COMMON_DIM = 512, 8 heads
    (17, 40, 2, 512, 4, False),      # hd = 128 (LIST_NUM_HEADS = 4)
    (9, 33, 1, 512, 2, False),       # hd = 256 (LIST_NUM_HEADS = 2)
    (197, 197, 2, 512, 8, True),     # ViT-sized self-attention; 394 rows -> projections on the tcgen05 engine
    (64, 300, 4, 256, 8, False),     # ragged key tiles, tcgen05 projections on the key side
    (1, 1, 32, 512, 8, False),       # the reference's own S = 1 call shape
    (3, 2, 1, 40, 5, False),         # hd = 8, D not a multiple of 8 -> FFMA projections
])
def test_fresh_vs_oracle(shape):
    Sq, Sk, B, D, H, sa = shape
    c = _fresh(Sq, Sk, B, D, H, sa, seed=Sq * 131 + Sk)
    ref = oracle_mha(c, H, sa)
    got = run_cuda(c, H, sa)
    for k, v in got.items():
        if np.abs(ref[k]).max() == 0:
            assert np.abs(v).max() < 1e-30, k
        else:
            assert parity.rel_err(v, ref[k]) < TOL, (k, parity.rel_err(v, ref[k]))


def test_state_dict_interchanges_with_torch_module():
    torch.manual_seed(3)
    ref = torch.nn.MultiheadAttention(128, 4).cuda()
    mine = fb.MultiheadAttention(128, 4).cuda()
    mine.load_state_dict(ref.state_dict(), strict=True)
    q = torch.randn(21, 3, 128, device="cuda"); kv = torch.randn(13, 3, 128, device="cuda")
    a, _ = ref(q, kv, kv)
    b, _ = mine(q, kv, kv)
    assert parity.rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-5


def test_rejects_bad_arguments():
    with pytest.raises(AssertionError):
        fb.MultiheadAttention(512, 7)
    m = fb.MultiheadAttention(64, 8).cuda()
    with pytest.raises(fb.Fb200Error):
        m(torch.randn(2, 1, 64), torch.randn(2, 1, 64), torch.randn(2, 1, 64))      # host tensors: no CPU path


@pytest.mark.parametrize("Sq,Sk,B,D,H", [(197, 85, 4, 128, 8), (7, 3, 5, 64, 4), (1, 1, 9, 64, 8)])
def test_mean_pool_folded_in_front_of_the_output_projection(Sq, Sk, B, D, H):
    """pool="mean" (multimodalGated.py:200-205: attention output averaged over its tokens) equals out.mean(0) of the plain
    module, forward and every gradient - the oracle is the float64 attention with the pooled gradient broadcast to the tokens."""
    c = _fresh(Sq, Sk, B, D, H, False, seed=Sq + B)
    dpool = np.random.default_rng(5).standard_normal((B, D))
    c["dy"] = np.broadcast_to(dpool[None] / Sq, (Sq, B, D)).copy()
    ref = oracle_mha(c, H, False)
    dev = "cuda"
    m = fb.MultiheadAttention(D, H, pool="mean").to(dev)
    with torch.no_grad():
        m.in_proj_weight.copy_(torch.from_numpy(c["in_w"])); m.in_proj_bias.copy_(torch.from_numpy(c["in_b"]))
        m.out_proj.weight.copy_(torch.from_numpy(c["out_w"])); m.out_proj.bias.copy_(torch.from_numpy(c["out_b"]))
    q, k, v = (torch.from_numpy(c[n]).float().to(dev).requires_grad_(True) for n in ("q", "k", "v"))
    out, _ = m(q, k, v)
    assert out.shape == (B, D)
    out.backward(torch.from_numpy(dpool).float().to(dev))
    torch.cuda.synchronize()
    got = dict(out=out, dq=q.grad, dk=k.grad, dv=v.grad, d_in_w=m.in_proj_weight.grad, d_in_b=m.in_proj_bias.grad,
               d_out_w=m.out_proj.weight.grad, d_out_b=m.out_proj.bias.grad)
    ref = dict(ref, out=ref["out"].mean(axis=0))
    for n, t in got.items():
        r = ref[n]
        if np.abs(r).max() == 0:
            assert np.all(t.detach().cpu().numpy() == 0), n
        else:
            assert parity.rel_err(t.detach().cpu().numpy(), r) < TOL, (n, parity.rel_err(t.detach().cpu().numpy(), r))
