"""debug: per-tensor errors of one golden case through the forced tcgen05 path"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-model-skin-lesion-classifier_b200"))
from tests import parity
from tests.golden import cases as _C; CASES = _C.all_cases()
from tests.gpu_util import build_model, run_autograd
from fusion_b200 import _lib
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_concat_eval"
flags = _lib.FLAG_FORCE_TC | (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
case = CASES[name]
for rep in range(2):
    cfg, model = build_model(case, "fp32", flags=flags)
    logits, loss, grads, dx = run_autograd(model, cfg, case)
    g = parity.load_golden(name)
    print("rep", rep, "logits", parity.rel_err(logits, g["logits64"]))
    for k in sorted(set(g["grad_names"].tolist())):
        got = np.asarray(grads[k], dtype=np.float64)
        if "g:" + k in g.files:
            e = parity.rel_err(got, g["g:" + k])
        else:
            maxabs, l2 = g["m:" + k]
            e = np.abs(got.ravel()[g["i:" + k]] - g["s:" + k]).max() / maxabs
        if e > 1e-5: print("   BAD", k, got.shape, f"{e:.3e}")
