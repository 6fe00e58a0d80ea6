"""Not a test: numpy emulation of an exactly-rounded bf16 pipeline (operands of every Linear rounded to bf16,
float64 elsewhere) against the float64 oracle, to quantify what ANY bf16 implementation can reach on the gradients
("bf16 parity").  Run: python tests/bf16_emulation.py"""
import sys; import os; sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import numpy as np
from oracle import head_oracle as ho
from tests.golden import cases as C
from tests import parity
def bf16(a):
    a = np.asarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).astype(np.float64)
# monkeypatch linear to quantise operands like the CUDA bf16 path: x, W in bf16 (fwd), dY bf16 (bwd)
def linear_q(t, x, W, b, qw=True):
    xq = bf16(x.v); Wq = bf16(W.v) if qw else W.v
    y = ho.Var(xq @ Wq.T + b.v)
    def bwd():
        if y.grad is None: return
        dyq = bf16(y.grad)
        x.acc(dyq @ Wq)
        g2 = dyq.reshape(-1, dyq.shape[-1])
        W.acc(g2.T @ xq.reshape(-1, xq.shape[-1]))
        b.acc(g2.sum(axis=0))
    t.push(bwd)
    return y
for qw in (False, True):
  ho.linear = lambda t,x,W,b: linear_q(t,x,W,b,qw)
  for name in ['cfg1_concat_eval','cfg2_cross_eval','cfg5_rgatt_eval', 'cfg3a_meta_eval']:
    case = C.all_cases()[name]; cfg = C.make_cfg(case['cfg'])
    params = C.gen_params(cfg, case['seed'], np.float64)
    x,tin,labels,cw,masks = C.gen_inputs(cfg, case['B'], case['seed'], case['train'], np.float64)
    o = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    g = parity.load_golden(name)
    errs = {}
    for k,got in o['grads'].items():
        if got is None: continue
        if 'm:'+k in g.files:
            maxabs,l2 = g['m:'+k]; samp = got.ravel()[g['i:'+k]]
            errs[k] = np.abs(samp-g['s:'+k]).max()/maxabs
    worst = sorted(errs.items(), key=lambda kv:-kv[1])[:4]
    print(qw, name, 'logits', parity.rel_err(o['logits'], g['logits64']), [(k, f'{v:.3f}') for k,v in worst])
print("---- exact vs emulated, per-tensor L2 and global")
import importlib
for name in ['cfg1_concat_eval','cfg2_cross_train','cfg5_rgatt_train', 'cfg3a_meta_eval', 'cfg4a_gfcam_train']:
    case = C.all_cases()[name]; cfg = C.make_cfg(case['cfg'])
    params = C.gen_params(cfg, case['seed'], np.float64)
    x,tin,labels,cw,masks = C.gen_inputs(cfg, case['B'], case['seed'], case['train'], np.float64)
    importlib.reload(ho)
    o0 = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    ho.linear = lambda t,x,W,b: linear_q(t,x,W,b,True)
    o1 = ho.head_forward_backward(cfg, params, x, tin, labels, cw, masks, need_input_grad=True)
    num=den=0; per=[]
    for k,g0 in o0['grads'].items():
        if g0 is None: continue
        d = o1['grads'][k]-g0
        num += (d*d).sum(); den += (g0*g0).sum()
        per.append((np.sqrt((d*d).sum()/(g0*g0).sum()), k))
    per.sort(reverse=True)
    print(name, 'global relL2 %.4f'%np.sqrt(num/den), 'logits relmax %.4f'%parity.rel_err(o1['logits'],o0['logits']), 'loss rel %.5f'%(abs(o1['loss']-o0['loss'])/o0['loss']), [(f'{a:.3f}',k) for a,k in per[:3]])
