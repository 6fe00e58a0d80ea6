import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/multimodal-model-skin-lesion-classifier_b200')
import numpy as np, torch
from oracle import head_oracle as ho
from tests import parity
from tests.golden import cases as C
from tests.gpu_util import build_model, run_autograd
from fusion_b200 import _lib
for B, flags, train in ((257,0,True),(257,0,False),(259,0,False),(260,0,False)):
    kw = dict(mechanism='crossattention', F=2048, V=85, C=6)
    case = dict(cfg=kw, B=B, seed=4242 + 257, train=train, full_grads=False)
    cfg, model = build_model(case, "fp32", flags=flags)
    logits, loss, grads, dx = run_autograd(model, cfg, case)
    params = C.gen_params(cfg, case["seed"], np.float64)
    x, tin, labels, cw, masks = C.gen_inputs(cfg, B, case["seed"], train, np.float64)
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    errs = {k: parity.rel_err(grads[k], g) for k,g in o['grads'].items() if g is not None}
    print(B, flags, train, 'logits %.2e'%parity.rel_err(logits,o['logits']), 'dx %.2e'%parity.rel_err(dx,o['d_img_feat']))
    for k in reversed(list(errs)): print('   %-45s %.2e'%(k, errs[k]))
    rowerr = np.abs(dx-o['d_img_feat']).max(1)/np.abs(o['d_img_feat']).max()
    print('   dx row errs > 1e-4 at rows', np.nonzero(rowerr>1e-4)[0][:20])
