"""GPU: primitive entry points of the C ABI against numpy restatements."""
import ctypes as Ct

import numpy as np
import pytest
import torch

import fusion_b200 as fb
from fusion_b200 import _lib
from oracle import head_oracle as ho
from tests import parity

pytestmark = pytest.mark.gpu


def vp(t):
    return Ct.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("B,Cn", [(1, 2), (7, 6), (32, 8), (1000, 6), (4096, 2), (33, 37)])
def test_cross_entropy(B, Cn):
    rng = np.random.default_rng(B * 100 + Cn)
    z = (3 * rng.standard_normal((B, Cn))).astype(np.float32)
    y = rng.integers(0, Cn, B)
    w = (rng.random(Cn) + 0.5).astype(np.float32)
    out, dl = fb.cross_entropy(torch.from_numpy(z).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(w).cuda())
    loss, dz, num, den = ho.weighted_cross_entropy(z.astype(np.float64), y, w.astype(np.float64))
    o = out.cpu().numpy()
    assert abs(o[0] - loss) < 1e-5 * abs(loss) and abs(o[1] - num) < 1e-5 * abs(num) and abs(o[2] - den) < 1e-5 * den
    assert parity.rel_err(dl.cpu().numpy(), dz) < 1e-5
    # unweighted + global denominator
    den2 = torch.tensor([den * 3.0], device="cuda", dtype=torch.float32)
    out2, dl2 = fb.cross_entropy(torch.from_numpy(z).cuda(), torch.from_numpy(y).cuda(), None, denom=den2)
    loss2, dz2, _, _ = ho.weighted_cross_entropy(z.astype(np.float64), y, None, denom=den * 3.0)
    assert parity.rel_err(dl2.cpu().numpy(), dz2) < 1e-5 and abs(out2.cpu().numpy()[0] - loss2) < 1e-5 * abs(loss2)


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(32, 512, 2048), (5, 6, 256), (33, 256, 85), (300, 130, 77), (1024, 512, 512), (128, 2048, 512)])
def test_simt_gemm_all_layouts(layout, M, N, K):
    rng = np.random.default_rng(M + N + K + layout)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    A = torch.from_numpy(a if layout != 2 else np.ascontiguousarray(a.T)).cuda()
    Bm = torch.from_numpy(np.ascontiguousarray(b.T) if layout == 0 else b).cuda()
    Cc = torch.full((M, N), 0.5, device="cuda")
    L = _lib.lib()
    _lib.check(L.fb200_gemm(layout, 0, M, N, K, vp(A), A.shape[1], vp(Bm), Bm.shape[1], vp(Cc), N, vp(torch.from_numpy(bias).cuda()), 1, 1, None, 0, None))
    torch.cuda.synchronize()
    ref = np.maximum(a.astype(np.float64) @ b.astype(np.float64) + bias, 0) + 0.5
    assert parity.rel_err(Cc.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("B,N,train", [(5, 64, True), (32, 512, True), (1000, 256, False), (3, 1664, True), (9, 2048, True)])
def test_ln_relu_dropout_fwd_bwd(B, N, train):
    rng = np.random.default_rng(B + N)
    x = rng.standard_normal((B, N)).astype(np.float32)
    g = (1 + 0.2 * rng.standard_normal(N)).astype(np.float32)
    b = (0.1 * rng.standard_normal(N)).astype(np.float32)
    mask = (rng.random((B, N)) >= 0.5).astype(np.uint8)
    dy = rng.standard_normal((B, N)).astype(np.float32)
    t = ho.Tape()
    X, G, Bv = ho.Var(x.astype(np.float64)), ho.Var(g.astype(np.float64)), ho.Var(b.astype(np.float64))
    Y = ho.dropout(t, ho.relu(t, ho.layernorm(t, X, G, Bv)), 0.5, mask if train else None)
    Y.grad = dy.astype(np.float64)
    t.backward()
    L = _lib.lib()
    dev = lambda a: torch.from_numpy(a).cuda()
    xs, gs, bs, ms, dys = dev(x), dev(g), dev(b), dev(mask), dev(dy)
    y = torch.empty_like(xs); st = torch.empty(B, 2, device="cuda"); dx = torch.empty_like(xs)
    dg = torch.empty(N, device="cuda"); db = torch.empty(N, device="cuda")
    _lib.check(L.fb200_ln_relu_dropout_fwd(vp(xs), vp(gs), vp(bs), vp(ms) if train else None, Ct.c_float(0.5), int(train), 0, 0, 4, B, N, vp(y), vp(st), None))
    _lib.check(L.fb200_ln_relu_dropout_bwd(vp(xs), vp(y), vp(gs), vp(st), vp(dys), Ct.c_float(0.5), int(train), B, N, vp(dx), vp(dg), vp(db), None))
    torch.cuda.synchronize()
    assert parity.rel_err(y.cpu().numpy(), Y.v) < 1e-5
    assert parity.rel_err(dx.cpu().numpy(), X.grad) < 1e-5
    assert parity.rel_err(dg.cpu().numpy(), G.grad) < 1e-5
    assert parity.rel_err(db.cpu().numpy(), Bv.grad) < 1e-5


@pytest.mark.parametrize("B,N", [(4, 48), (32, 1664), (7, 2048), (300, 768)])
def test_metablock_fwd(B, N):
    rng = np.random.default_rng(B * N)
    mk = lambda *s: rng.standard_normal(s).astype(np.float32)
    v, f, g = mk(B, N), mk(B, N), mk(B, N)
    gf, bf, gg, bg = 1 + 0.2 * mk(N), 0.1 * mk(N), 1 + 0.2 * mk(N), 0.1 * mk(N)
    def ln(x, gm, bt):
        x = x.astype(np.float64); mu = x.mean(1, keepdims=True); var = ((x - mu) ** 2).mean(1, keepdims=True)
        return (x - mu) / np.sqrt(var + 1e-5) * gm + bt
    ref = 1 / (1 + np.exp(-(np.tanh(v * ln(f, gf, bf)) + ln(g, gg, bg))))
    L = _lib.lib()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    ts = [dev(a) for a in (v, f, g, gf, bf, gg, bg)]
    y = torch.empty(B, N, device="cuda"); st = torch.empty(B, 4, device="cuda")
    _lib.check(L.fb200_metablock_fwd(*[vp(t) for t in ts], B, N, vp(y), vp(st), None))
    torch.cuda.synchronize()
    assert parity.rel_err(y.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 512), (1024, 512, 2048), (512, 2048, 1024), (200, 136, 72), (40, 512, 512)])
def test_tcgen05_gemm_all_layouts(engine, layout, M, N, K):
    """tcgen05 GEMM (engine 1 = 3xTF32 fp32-strict, 2 = bf16) on K-major and MN-major operands,
    ragged tiles included, with bias + ReLU + accumulate in the epilogue."""
    rng = np.random.default_rng(M + N + K + layout)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    A = torch.from_numpy(a if layout != 2 else np.ascontiguousarray(a.T)).cuda()
    Bm = torch.from_numpy(np.ascontiguousarray(b.T) if layout == 0 else b).cuda()
    Cc = torch.full((M, N), 0.5, device="cuda")
    L = _lib.lib()
    wsz = Ct.c_size_t(0)
    _lib.check(L.fb200_gemm_workspace_bytes(layout, engine, M, N, K, Ct.byref(wsz)))
    ws = torch.empty(wsz.value, dtype=torch.uint8, device="cuda")
    _lib.check(L.fb200_gemm(layout, engine, M, N, K, vp(A), A.shape[1], vp(Bm), Bm.shape[1], vp(Cc), N, vp(torch.from_numpy(bias).cuda()), 1, 1,
                            vp(ws), ws.numel(), None))
    torch.cuda.synchronize()
    if engine == 2:
        q = lambda x: torch.from_numpy(x).bfloat16().double().numpy()
        ref = q(a) @ q(b)
    else:
        ref = a.astype(np.float64) @ b.astype(np.float64)
    ref = np.maximum(ref + bias, 0) + 0.5
    assert parity.rel_err(Cc.cpu().numpy(), ref) < (3e-6 if engine == 1 else 1e-5)


def test_tcgen05_split_k_weight_gradient():
    """dW-shaped TN GEMM with the reduction over a long batch: split-K + fp32 atomics."""
    M, N, K = 512, 512, 8192
    rng = np.random.default_rng(3)
    a = rng.standard_normal((K, M)).astype(np.float32); b = rng.standard_normal((K, N)).astype(np.float32)
    A, Bm = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    Cc = torch.empty(M, N, device="cuda")
    L = _lib.lib(); wsz = Ct.c_size_t(0)
    _lib.check(L.fb200_gemm_workspace_bytes(2, 1, M, N, K, Ct.byref(wsz)))
    ws = torch.empty(wsz.value, dtype=torch.uint8, device="cuda")
    _lib.check(L.fb200_gemm(2, 1, M, N, K, vp(A), M, vp(Bm), N, vp(Cc), N, None, 0, 0, vp(ws), ws.numel(), None))
    torch.cuda.synchronize()
    assert parity.rel_err(Cc.cpu().numpy(), a.astype(np.float64).T @ b.astype(np.float64)) < 3e-6


def test_cross_entropy_ignores_out_of_range_labels():
    """ignore_index semantics (-100, or any label outside [0, C)): no contribution, never an out-of-range read."""
    rng = np.random.default_rng(5)
    z = rng.standard_normal((64, 6)).astype(np.float32)
    y = rng.integers(0, 6, 64); y[3] = -100; y[10] = 6; y[11] = 2 ** 40
    w = (rng.random(6) + 0.5).astype(np.float32)
    out, dl = fb.cross_entropy(torch.from_numpy(z).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(w).cuda())
    loss, dz, num, den = ho.weighted_cross_entropy(z.astype(np.float64), y, w.astype(np.float64))
    assert abs(out.cpu().numpy()[0] - loss) < 1e-5 * abs(loss)
    assert parity.rel_err(dl.cpu().numpy(), dz) < 1e-5 and (dl.cpu().numpy()[[3, 10, 11]] == 0).all()
    ref = torch.nn.functional.cross_entropy(torch.from_numpy(z), torch.from_numpy(np.where((y >= 0) & (y < 6), y, -100)), weight=torch.from_numpy(w))
    assert abs(float(ref) - loss) < 1e-5
