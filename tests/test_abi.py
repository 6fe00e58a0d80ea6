"""CPU: the C-ABI library loads, exports every symbol include/fb200.h declares, and its
host-side introspection (slots, shapes, gradient liveness, dropout sites, work model) agrees
with the oracle and with the reference's None/zero gradient pattern stored in the goldens.
No compute entry point is called here (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from fusion_b200 import _lib, make_desc
from oracle import head_oracle as ho
from tests import parity
from tests.golden import cases as C

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"libfb200.so does not export {n}"
    assert _lib.lib().fb200_version() == 100


def test_strerror_and_mechanism_strings():
    L = _lib.lib()
    for i, m in enumerate(ho.MECHANISMS):
        assert L.fb200_mechanism_from_string(m.encode()) == i
        assert L.fb200_mechanism_string(i).decode() == m
    assert L.fb200_mechanism_from_string(b"metablock-se") == -1        # referenced at :114 but never implemented
    assert b"no CPU path" in L.fb200_strerror(-2)


@pytest.mark.parametrize("mech", ho.MECHANISMS)
def test_slots_match_oracle_shapes(mech):
    cfg = ho.HeadConfig(mech, F=2048, C=6, V=85)
    d = make_desc(mech, 32, 2048, 85, 512, 512, 8, 6)
    shapes = cfg.param_shapes()
    names = _lib.param_names()
    assert names == list(shapes.keys())
    for i, k in enumerate(names):
        assert _lib.param_shape(d, i) == tuple(shapes[k]), k


def test_text_mode1_has_no_text_fc():
    d = make_desc("gfcam", 8, 768, 0, 85, 512, 8, 2, text_mode=1)
    names = _lib.param_names()
    for i, k in enumerate(names):
        shp = _lib.param_shape(d, i)
        assert (shp is None) == k.startswith("text_fc."), k
    assert _lib.param_shape(d, names.index("text_projector.weight")) == (512, 85)


@pytest.mark.parametrize("name", sorted(n for n in C.all_cases() if n.startswith("small") or n.endswith("_eval")))
def test_gradient_liveness_matches_reference(name):
    """fb200_grad_offset >= 0 exactly for the parameters whose .grad is a tensor in the reference."""
    case = C.all_cases()[name]
    kw = case["cfg"]
    g = parity.load_golden(name)
    d = make_desc(kw["mechanism"], case["B"], kw["F"], kw.get("V") or 0, kw.get("T", 512), kw.get("D", 512), kw.get("H", 8), kw["C"],
                  text_mode=0 if kw.get("text_model", "one-hot-encoder") == "one-hot-encoder" else 1)
    total, offs = _lib.grad_layout(d)
    names = _lib.param_names()
    live = {names[s] for s in offs}
    assert live == set(g["grad_names"].tolist())
    assert not (live & set(g["none_grads"].tolist()))
    # offsets are 16-byte aligned, disjoint and inside the flat buffer
    spans = sorted((o, int(np.prod(_lib.param_shape(d, s)))) for s, o in offs.items())
    for (o, n), (o2, _) in zip(spans, spans[1:] + [(total, 0)]):
        assert o % 4 == 0 and o + n <= o2


def test_work_model_matches_survey_table():
    """live params and algorithmic FLOPs/bytes of the BASELINE configs."""
    rows = {  # mechanism, F, V, C, T, text_mode -> live params (M), bytes (MB)
        "cfg1": (("concatenation", 512, 85, 6, 512, 0), 1.601e6, 19.4e6),
        "cfg2": (("crossattention", 2048, 85, 6, 512, 0), 4.488e6, 54.4e6),
        "cfg3a": (("metablock", 1664, 13, 8, 512, 0), 3.099e6, 37.6e6),
        "cfg3b": (("weighted", 1664, 13, 8, 512, 0), 2.698e6, 32.8e6),
        "cfg5": ((ho.RG_ATT, 1024, 85, 6, 512, 0), 5.542e6, None),
    }
    for k, ((m, F, V, Cn, T, tm), plive, nbytes) in rows.items():
        d = make_desc(m, 32, F, V, T, 512, 8, Cn, text_mode=tm, flags=_lib.FLAG_NEED_DIMG)
        flops, b, p = _lib.algorithmic_work(d)
        assert abs(p - plive) / plive < 2e-3, (k, p)
        if nbytes:
            assert abs(b - nbytes) / nbytes < 1e-2, (k, b)
        assert abs(flops - 6 * 32 * p) / flops < 0.03, (k, flops)    # 6*B*MAC_live minus the text_fc.0 dX


def test_bad_descriptors_are_rejected():
    L = _lib.lib()
    n = ctypes.c_size_t()
    d = make_desc("crossattention", 32, 2048, 85, 512, 512, 7, 6)          # D % H != 0 -> like nn.MultiheadAttention's assert
    assert L.fb200_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == -1
    d = make_desc("crossattention", 0, 2048, 85, 512, 512, 8, 6)
    assert L.fb200_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == -1
    d = make_desc("crossattention", 32, 2048, 85, 512, 512, 8, 6, n=1)
    assert L.fb200_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == -1
    with pytest.raises(ValueError, match="not implemented"):
        make_desc("metablock-se", 32, 2048, 85, 512, 512, 8, 6)


def test_dropout_sites():
    d = make_desc(ho.RG_ATT, 16, 1024, 85, 512, 512, 8, 6, train=True)
    s = _lib.dropout_sites(d)
    assert s == {0: (pytest.approx(0.1), 16, 512), 1: (pytest.approx(0.1), 16, 512),
                 4: (pytest.approx(0.5), 16, 512), 5: (pytest.approx(0.5), 16, 256)}
    d = make_desc("metablock", 16, 1664, 13, 512, 512, 8, 8, train=True)
    s = _lib.dropout_sites(d)
    assert set(s) == {4, 5} and abs(s[4][0] - 0.3) < 1e-7


def test_dp_bucket_split_is_a_balanced_weight_boundary():
    """fb200_dp_bucket_split (host-only): 0 where nothing can be split (FFMA path), otherwise the offset of a weight
    gradient that cuts the tcgen05 weight-gradient tiles into two halves of equal wave count for the headline config."""
    L = _lib.lib()
    small = make_desc("crossattention", 32, 2048, 85, 512, 512, 8, 6, train=True)
    assert _lib.dp_bucket_split(small) == 0
    d = make_desc("crossattention", 4096, 2048, 85, 512, 512, 8, 6, train=True)
    split = _lib.dp_bucket_split(d)
    total, offs = _lib.grad_layout(d)
    assert 0 < split < total
    names = _lib.param_names()
    # the split is the start of a weight (or of the V third of an in_proj_weight): never inside a bias / LayerNorm vector
    starts = set()
    for i, k in enumerate(names):
        if i in offs and k.endswith("weight"):
            shp = _lib.param_shape(d, i)
            starts.add(offs[i])
            if len(shp) == 2 and k.endswith("in_proj_weight"):
                starts.add(offs[i] + 2 * shp[0] // 3 * shp[1])
    assert split in starts
    # tiles (128 x 128) of the tcgen05 weight gradients on either side: 136 | 136 for this configuration = one wave each on 148 SMs
    lo = hi = 0
    for i, k in enumerate(names):
        if i not in offs or not k.endswith("weight"):
            continue
        shp = _lib.param_shape(d, i)
        if len(shp) != 2:
            continue                                      # LayerNorm weight vectors
        r, c = shp
        if c % 8 or c < 32 or r < 32:
            continue                                      # K = 85 (FFMA), the 6-class head
        if k.endswith("in_proj_weight"):
            r, off = r // 3, offs[i] + 2 * (r // 3) * c   # only the V third is computed
        else:
            off = offs[i]
        tiles = -(-r // 128) * -(-c // 128)
        if off < split:
            lo += tiles
        else:
            hi += tiles
    assert (lo, hi) == (136, 136)


def test_mha_workspace_and_argument_checks_on_host():
    L = _lib.lib()
    d = _lib.MhaDesc(Sq=197, Skv=85, B=32, D=512, H=8, flags=0)
    n = ctypes.c_size_t(0)
    assert L.fb200_mha_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == 0
    floats = 6 * 197 * 32 * 512 + 4 * 85 * 32 * 512 + 2 * 32 * 8 * 197        # Q, O, dO, dQ | K, V, dK, dV | lse, delta ... (+ alignment)
    assert n.value >= 4 * (4 * 197 * 32 * 512 + 4 * 85 * 32 * 512 + 2 * 32 * 8 * 197) and n.value < 4 * floats
    bad = _lib.MhaDesc(Sq=4, Skv=4, B=1, D=512, H=7, flags=0)                   # embed_dim % num_heads != 0
    assert L.fb200_mha_workspace_bytes(ctypes.byref(bad), ctypes.byref(n)) == -1
    wide = _lib.MhaDesc(Sq=4, Skv=4, B=1, D=1024, H=2, flags=0)                 # head dim 512 > 256
    assert L.fb200_mha_workspace_bytes(ctypes.byref(wide), ctypes.byref(n)) == -2


def test_engine_policy_and_launch_counts_on_host():
    """Host-side plan introspection: fp32 uses the exact FFMA kernels up to 32 rows and the tcgen05 3xTF32 path above
    (bf16 always tcgen05); widths that TMA cannot take (V = 85) and the class head stay off the tensor path; the step of the
    headline configuration is 40 kernel launches (27 tcgen05 GEMMs + ONE grouped weight-gradient launch + 12 others)."""
    L = _lib.lib()

    def gemms(B, dtype="fp32", flags=0):
        d = make_desc("crossattention", B, 2048, 85, 512, 512, 8, 6, dtype=dtype, train=True, flags=flags)
        arr = (ctypes.c_int32 * (5 * 256))()
        L.fb200_list_gemms.restype = ctypes.c_int
        L.fb200_list_gemms.argtypes = [ctypes.POINTER(_lib.Desc), ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        n = L.fb200_list_gemms(ctypes.byref(d), arr, 256)
        return [tuple(arr[5 * i + j] for j in range(5)) for i in range(n)], d      # (layout, engine, M, N, K)

    g32, _ = gemms(32)
    assert {e for _, e, *_ in g32} == {0}                                          # FFMA everywhere
    assert {e for _, e, *_ in gemms(64, flags=_lib.FLAG_FORCE_MEGA)[0]} == {0}    # persistent step kernel (up to 64 rows when forced): FFMA arithmetic
    assert gemms(64)[0] == gemms(64, flags=_lib.FLAG_NO_MEGA)[0]                  # default above 32 rows: tcgen05 with cluster split-K
    g64, _ = gemms(64, flags=_lib.FLAG_NO_MEGA)
    assert {e for _, e, *_ in g64} == {0, 1}
    assert all(e == 0 for lay, e, M, N, K in g64 if 85 in (N, K))                  # text_fc.0: K = 85 is not TMA-legal
    assert all(e == 1 for lay, e, M, N, K in g64 if 85 not in (N, K))
    gb, _ = gemms(32, dtype="bf16")
    assert 2 in {e for _, e, *_ in gb}                                             # bf16 rides tcgen05 at any batch
    gs, _ = gemms(4096, flags=_lib.FLAG_FORCE_SIMT)
    assert {e for _, e, *_ in gs} == {0}
    g, d = gemms(4096)
    fwd, bwd = ctypes.c_int(0), ctypes.c_int(0)
    assert L.fb200_launch_count(ctypes.byref(d), ctypes.byref(fwd), ctypes.byref(bwd)) == 0
    assert (fwd.value, bwd.value) == (18, 19)              # + CE pass 1 / 2 + the Philox advance = the 40 launches of profiles/r01_launch_list
    # forward NT, weight-gradient TN for every Linear; input-gradient NN only where something upstream needs it
    nt = [x for x in g if x[0] == 0]; tn = [x for x in g if x[0] == 2]; nn = [x for x in g if x[0] == 1]
    assert len(nt) == len(tn) == 15 and len(nn) == 13      # image_projector and text_fc.0 have no dX (inputs need no gradient)


def test_step_kernel_program_fits_for_every_fusion_string():
    """Host only: the persistent step kernel's program (csrc/mega.cuh) for all 18 fusion strings, edge batches and both input-gradient
    flags stays inside the kernel-parameter capacity (128 GEMM ops, 26 row ops, 80 stages), chains the classifier tail into one
    task, and batches above 64 rows / bf16 / forced engines take the per-op kernels."""
    from oracle import head_oracle as ho
    worst = [0, 0, 0]
    for mech in ho.MECHANISMS:
        for B in (1, 32, 33, 64):
            for flags in (0, _lib.FLAG_NEED_DIMG | _lib.FLAG_NEED_DTEXT):
                d = make_desc(mech, B, 2048, 85, 512, 512, 8, 6, n=2, train=True, flags=flags | _lib.FLAG_FORCE_MEGA)
                for which in (0, 1, 2):
                    info = _lib.mega_program_info(d, which)
                    assert info is not None, (mech, B, which)
                    stages, ng, nr, tasks = info
                    assert 1 <= stages <= 80 and ng <= 128 and nr <= 26 and tasks >= 1, (mech, B, which, info)
                    worst = [max(a, b) for a, b in zip(worst, info[:3])]
    print("largest program: stages, gemm ops, row ops =", worst)
    d = make_desc("crossattention", 32, 2048, 85, 512, 512, 8, 6, train=True)
    fwd, bwd, step = (_lib.mega_program_info(d, w) for w in (0, 1, 2))
    assert step[0] < fwd[0] + bwd[0]                     # the tail (LayerNorm -> head -> CE -> their backward) shares one stage
    assert step[1] == fwd[1] + bwd[1]
    assert _lib.mega_program_info(make_desc("crossattention", 33, 2048, 85, 512, 512, 8, 6, train=True)) is None       # default: up to 32 rows
    for kw in (dict(B=65, flags=_lib.FLAG_FORCE_MEGA), dict(B=32, dtype="bf16"), dict(B=32, flags=_lib.FLAG_FORCE_SIMT), dict(B=32, flags=_lib.FLAG_NO_MEGA)):
        B = kw.pop("B")
        assert _lib.mega_program_info(make_desc("crossattention", B, 2048, 85, 512, 512, 8, 6, **kw)) is None
