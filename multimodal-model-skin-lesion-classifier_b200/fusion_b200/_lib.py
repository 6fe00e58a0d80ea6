"""ctypes binding of libfb200.so (the C ABI declared in include/fb200.h).

The library is built in-tree by ``multimodal-model-skin-lesion-classifier_b200/build.sh``
(``__graft_entry__.build()``).  There is no Python or CPU fallback: if the shared object
is missing, importing this module raises, and every compute entry point refuses host
pointers with FB200_EUNSUPPORTED.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FB200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libfb200.so")   # FB200_LIB: A/B builds of the same library

F32, BF16 = 0, 1
FLAG_NEED_DIMG, FLAG_NEED_DTEXT, FLAG_FORCE_SIMT, FLAG_FORCE_TC, FLAG_ONE_STREAM, FLAG_NO_MEGA, FLAG_FORCE_MEGA = 1, 2, 4, 8, 16, 32, 64
NUM_DROPOUT_SITES = 6
DROP_SITES = ("img_res", "txt_res", "img_res2", "txt_res2", "fc1", "fc2")


class Desc(C.Structure):
    """struct fb200_desc (include/fb200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "mechanism", "B", "F", "V", "T", "D", "H", "C", "n", "text_mode", "dtype", "train", "flags", "reserved")]


class MhaDesc(C.Structure):
    """struct fb200_mha_desc (include/fb200.h)."""
    _fields_ = [(n, C.c_int32) for n in ("Sq", "Skv", "B", "D", "H", "flags")]


class TabtDesc(C.Structure):
    """struct fb200_tabt_desc (include/fb200.h)."""
    _fields_ = [(n, C.c_int32) for n in ("B", "T", "D", "H", "F", "L", "n_emb_rows", "train")] + [("p", C.c_float), ("flags", C.c_int32)]


class Fb200Error(RuntimeError):
    """Non-zero status from libfb200 (kept a RuntimeError so the reference's
    ``try/except ... continue`` around each experiment keeps working: train_pad_20.py:486-488)."""

    def __init__(self, status, what=""):
        self.status = status
        msg = lib().fb200_strerror(status).decode()
        super().__init__(f"{msg} [{status}]" + (f" in {what}" if what else ""))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with multimodal-model-skin-lesion-classifier_b200/build.sh "
            "(python -c 'import __graft_entry__ as g; g.build()').  fusion_b200 has no fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_size_t
    dp = C.POINTER(Desc)
    pp = C.POINTER(C.c_void_p)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype, f.argtypes = res, list(args)

    sig("fb200_version", i32)
    sig("fb200_strerror", C.c_char_p, i32)
    sig("fb200_mechanism_from_string", i32, C.c_char_p)
    sig("fb200_mechanism_string", C.c_char_p, i32)
    sig("fb200_num_params", i32)
    sig("fb200_param_name", C.c_char_p, i32)
    sig("fb200_param_shape", i32, dp, i32, C.POINTER(i64), C.POINTER(i64))
    sig("fb200_grad_offset", i64, dp, i32)
    sig("fb200_grad_elems", i64, dp)
    sig("fb200_workspace_bytes", i32, dp, C.POINTER(sz))
    sig("fb200_grad_live_ranges", i32, dp, C.POINTER(i64), i32)
    sig("fb200_dropout_p", C.c_float, dp, i32)
    sig("fb200_dropout_shape", i32, dp, i32, C.POINTER(i64), C.POINTER(i64))
    sig("fb200_algorithmic_work", i32, dp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64))
    sig("fb200_launch_count", i32, dp, C.POINTER(i32), C.POINTER(i32))
    sig("fb200_head_forward", i32, dp, pp, vp, vp, pp, u64, u64, vp, vp, vp, vp)
    sig("fb200_rng_advance", i32, vp, u64, vp)
    sig("fb200_aux_loss", i32, i32, vp, vp, vp, C.c_float, i32, i32, vp, vp, vp)
    sig("fb200_softmax_argmax", i32, vp, i32, i32, vp, vp, vp)
    sig("fb200_metadata_encode", i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp)
    sig("fb200_adam_step", i32, i32, pp, pp, pp, pp, C.POINTER(i64), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i64, C.c_float, vp)
    sig("fb200_debug_gemm_replay", i32, dp, pp, vp, vp, vp, vp, vp, vp)
    sig("fb200_debug_tc_trace", i32, vp)
    sig("fb200_debug_tc_timeline", i32, vp, i32)
    sig("fb200_debug_set_pdl", i32, i32)
    sig("fb200_debug_mega_trace", i32, vp)
    sig("fb200_mega_program_info", i32, dp, i32, C.POINTER(i32))
    sig("fb200_debug_mega_barriers", i32, i32, vp, vp)
    mp = C.POINTER(MhaDesc)
    sig("fb200_mha_workspace_bytes", i32, mp, C.POINTER(sz))
    sig("fb200_mha_forward", i32, mp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp)
    sig("fb200_mha_backward", i32, mp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp)
    sig("fb200_list_gemms", i32, dp, C.POINTER(C.c_int32), i32)
    sig("fb200_head_backward", i32, dp, pp, vp, vp, pp, u64, u64, vp, vp, vp, vp, vp, vp, vp)
    sig("fb200_cross_entropy", i32, vp, vp, vp, vp, i32, i32, vp, vp, vp)
    sig("fb200_head_train_step", i32, dp, pp, vp, vp, vp, vp, vp, pp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp)
    sig("fb200_head_train_step_dp", i32, dp, pp, vp, vp, vp, vp, vp, pp, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp)
    sig("fb200_dp_bucket_split", i64, dp)
    sig("fb200_dp_allreduce", i32, vp, pp, C.POINTER(i64), i32, i32, i32, i32, pp, vp, i32, vp)
    f32 = C.c_float
    sig("fb200_gemm", i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp, i32, i32, vp, sz, vp)
    sig("fb200_gemm_workspace_bytes", i32, i32, i32, i32, i32, i32, C.POINTER(sz))
    sig("fb200_ln_relu_dropout_fwd", i32, vp, vp, vp, vp, f32, i32, u64, u64, i32, i32, i32, vp, vp, vp)
    sig("fb200_ln_relu_dropout_bwd", i32, vp, vp, vp, vp, vp, f32, i32, i32, i32, vp, vp, vp, vp)
    sig("fb200_metablock_fwd", i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp)
    tp = C.POINTER(TabtDesc)
    sig("fb200_tabt_param_elems", i32, tp, C.POINTER(i64), C.POINTER(i64))
    sig("fb200_tabt_workspace_bytes", i32, tp, C.POINTER(sz), C.POINTER(sz))
    sig("fb200_tabt_forward", i32, tp, vp, vp, vp, pp, u64, u64, vp, vp, i32, vp, vp)
    sig("fb200_tabt_backward", i32, tp, vp, vp, vp, pp, u64, u64, vp, vp, vp, i32, vp, vp, vp)
    sig("fb200_debug_tabt_trace", i32, vp)
    sig("fb200_linear_forward", i32, i32, i32, i32, vp, i32, vp, vp, i32, vp, i32, vp)
    sig("fb200_linear_backward", i32, i32, i32, i32, vp, i32, vp, vp, i32, vp, i32, vp, vp, vp)
    _lib = L
    return L


def check(status, what=""):
    if status != 0:
        raise Fb200Error(status, what)


def param_names():
    L = lib()
    return [L.fb200_param_name(i).decode() for i in range(L.fb200_num_params())]


def mechanism_id(name: str) -> int:
    return lib().fb200_mechanism_from_string(name.encode())


def param_shape(desc: Desc, slot: int):
    """Shape tuple of a slot under ``desc`` or None when the slot is absent."""
    r, c = C.c_int64(), C.c_int64()
    st = lib().fb200_param_shape(C.byref(desc), slot, C.byref(r), C.byref(c))
    if st == -5:
        return None
    check(st, "fb200_param_shape")
    return (r.value, c.value) if c.value else (r.value,)


_FIELDS = [f for f, _ in Desc._fields_]
_ws_cache, _layout_cache = {}, {}


def desc_key(desc: Desc):
    return tuple(getattr(desc, f) for f in _FIELDS)


def workspace_bytes(desc: Desc) -> int:
    key = desc_key(desc)
    hit = _ws_cache.get(key)
    if hit is not None:
        return hit
    n = C.c_size_t()
    check(lib().fb200_workspace_bytes(C.byref(desc), C.byref(n)), "fb200_workspace_bytes")
    _ws_cache[key] = n.value
    return n.value


def grad_layout(desc: Desc):
    """(total elements, {slot: offset}) of the flat gradient buffer for live slots (cached per descriptor)."""
    key = desc_key(desc)[:10] + (0, 0, 0, 0)          # the layout depends on the model shape only
    hit = _layout_cache.get(key)
    if hit is not None:
        return hit
    r = _grad_layout_uncached(desc)
    _layout_cache[key] = r
    return r


def _grad_layout_uncached(desc: Desc):
    L = lib()
    total = L.fb200_grad_elems(C.byref(desc))
    if total < 0:
        raise Fb200Error(-1, "fb200_grad_elems")
    offs = {}
    for s in range(L.fb200_num_params()):
        o = L.fb200_grad_offset(C.byref(desc), s)
        if o >= 0:
            offs[s] = o
    return total, offs


def algorithmic_work(desc: Desc):
    f, b, p = C.c_double(), C.c_double(), C.c_int64()
    check(lib().fb200_algorithmic_work(C.byref(desc), C.byref(f), C.byref(b), C.byref(p)), "fb200_algorithmic_work")
    return f.value, b.value, p.value


def launch_count(desc: Desc):
    f, b = C.c_int(), C.c_int()
    check(lib().fb200_launch_count(C.byref(desc), C.byref(f), C.byref(b)), "fb200_launch_count")
    return f.value, b.value


def mega_program_info(desc: Desc, which=2):
    """(stages, gemm ops, row ops, tile tasks) of the persistent step kernel's program for `desc` (host only), or None when
    the descriptor takes the per-op kernels.  which: 0 forward, 1 backward, 2 fused train step."""
    out = (C.c_int * 4)()
    st = lib().fb200_mega_program_info(C.byref(desc), which, out)
    if st == -2:
        return None
    check(st, "fb200_mega_program_info")
    return tuple(out)


def dropout_sites(desc: Desc):
    """{site index: (p, rows, cols)} for the dropout sites the mechanism reaches."""
    L = lib()
    out = {}
    for s in range(NUM_DROPOUT_SITES):
        p = L.fb200_dropout_p(C.byref(desc), s)
        if p > 0:
            r, c = C.c_int64(), C.c_int64()
            check(L.fb200_dropout_shape(C.byref(desc), s, C.byref(r), C.byref(c)), "fb200_dropout_shape")
            out[s] = (p, r.value, c.value)
    return out


_ranges_cache = {}


def dp_bucket_split(desc: Desc):
    """Element offset in the flat gradient buffer below which every gradient is final when the mid-step event of
    fb200_head_train_step_dp fires (0: nothing to split)."""
    v = lib().fb200_dp_bucket_split(C.byref(desc))
    if v < 0:
        raise Fb200Error(-1, "fb200_dp_bucket_split")
    return int(v)


def grad_live_ranges(desc: Desc):
    """[(begin, end)] element ranges of the flat gradient buffer that can be non-zero."""
    key = desc_key(desc)[:10]
    hit = _ranges_cache.get(key)
    if hit is not None:
        return hit
    cap = 64
    arr = (C.c_int64 * (2 * cap))()
    n = lib().fb200_grad_live_ranges(C.byref(desc), arr, cap)
    if n < 0:
        raise Fb200Error(n, "fb200_grad_live_ranges")
    out = [(arr[2 * i], arr[2 * i + 1]) for i in range(min(n, cap))]
    _ranges_cache[key] = out
    return out
