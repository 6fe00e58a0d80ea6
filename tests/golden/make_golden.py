"""Generate the golden fixtures from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py

For every case in cases.py the reference ``MultimodalModel`` (imported through
oracle/ref_shim.py) is run on CPU in float64 (tie-breaker precision for the 1e-5 check)
and in float32 (the precision the reference ships), forward + weighted CE + backward, with
the dropout keep-masks of the case injected.  Stored per case (npz):
  logits64, loss64, logits32, loss32, none_grads (names whose .grad stays None),
  small cases : every gradient of <= FULL_GRAD_MAX_ELEMS elements in full (float32 storage);
  otherwise   : per-gradient summary (seeded projections, sampled entries, max-abs, l2).
Inputs/weights are NOT stored: cases.py regenerates them from the seed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_shim  # noqa: E402
from tests.golden import cases as C  # noqa: E402


def run_reference(case, np_dtype):
    cfg = C.make_cfg(case["cfg"])
    B = case["B"]
    params = C.gen_params(cfg, case["seed"], np_dtype)
    x, tin, labels, cw, masks = C.gen_inputs(cfg, B, case["seed"], case["train"], np_dtype)
    tdt = torch.float64 if np_dtype == np.float64 else torch.float32
    model = ref_shim.build_reference_model(cfg.mechanism, cfg.F, cfg.C, V=cfg.V, T=cfg.T, D=cfg.D, H=cfg.H,
                                           text_model=cfg.text_model).to(tdt)
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    assert all(k.startswith(("text_encoder.", "image_encoder.")) for k in missing.missing_keys), missing
    assert not missing.unexpected_keys, missing
    model.train(case["train"])
    xt = torch.from_numpy(x).requires_grad_(True)
    tt = torch.from_numpy(tin)
    mq = [] if masks is None else [torch.from_numpy(masks[k]) for k in C.mask_keys(cfg)]
    with ref_shim.injected_dropout(mq) as q:
        logits = ref_shim.reference_forward(model, xt, tt)
        assert not q, "unused dropout masks: the case's mask list does not match the reference"
    loss = torch.nn.CrossEntropyLoss(weight=torch.from_numpy(cw))(logits, torch.from_numpy(labels))
    loss.backward()
    grads = {}
    for k, p in model.named_parameters():
        if k.startswith(("text_encoder.", "image_encoder.")):
            continue
        grads[k] = None if p.grad is None else p.grad.detach().numpy().copy()
    return logits.detach().numpy().copy(), float(loss), grads, xt.grad.numpy().copy()


def main():
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, case in C.all_cases().items():
        if only and name not in only:
            continue
        l64, loss64, g64, dx64 = run_reference(case, np.float64)
        l32, loss32, _, _ = run_reference(case, np.float32)
        out = dict(logits64=l64, loss64=np.float64(loss64), logits32=l32, loss32=np.float32(loss32),
                   none_grads=np.array(sorted(k for k, g in g64.items() if g is None)),
                   grad_names=np.array(sorted(k for k, g in g64.items() if g is not None)))
        def put_summary(k, g):
            s = C.grad_summary(k, g, case["seed"])
            out["p:" + k], out["i:" + k], out["s:" + k] = s["probes"], s["idx"], s["samples"]
            out["m:" + k] = np.array([s["maxabs"], s["l2"]])

        for k, g in list(g64.items()) + [("d_img_feat", dx64)]:
            if g is None:
                continue
            if case["full_grads"] and g.size <= C.FULL_GRAD_MAX_ELEMS:
                out["g:" + k] = g.astype(np.float32)      # 6e-8 relative: far below the 1e-5 bar
            else:
                put_summary(k, g)
        path = os.path.join(C.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:36s} {os.path.getsize(path) / 1024:8.1f} KiB  loss64={loss64:.6f}")


if __name__ == "__main__":
    main()
