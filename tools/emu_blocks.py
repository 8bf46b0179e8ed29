#!/usr/bin/env python
"""GPU diagnostic: native step vs the oracle's bf16 emulation (and vs fp32), block by block, forward and backward.
    python tools/emu_blocks.py [fixture name]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from oracle import msgchn_oracle as O
from golden_util import load_golden, case_frame, case_checkpoint, nrel, W_SD, W_SM, W_COS
from oracle_trace import trace_step, to_nchw
from test_msgchn_step_gpu import make_model, FWD_NAMES, GRAD_NAMES, ZERO_GRAD

name = sys.argv[1] if len(sys.argv) > 1 else 'msgchn_fit_kitti_1x64x128'
fx = load_golden(name); case = fx['case']
sd = case_checkpoint(case)
model = make_model(case, sd, case['max_input_depth'])
image, sparse, _ = case_frame(case, 0)
res = {}
for tag, pr in (('fp32', O.FP32), ('bf16', O.Precision('bf16'))):
    res[tag] = trace_step({k: v.clone() for k, v in sd.items()}, image, sparse, case['max_input_depth'], W_SD, W_SM, W_COS, pr=pr)
eng = model.model._engine_for(image.cuda())
eng.set_adam(0.0)
model.tta_step(image.cuda(), sparse.cuda(), 0.0, W_SD, W_SM, W_COS)
torch.cuda.synchronize()
print('%-14s %12s %12s %12s' % ('block', 'nat-vs-fp32', 'nat-vs-emu', 'emu-vs-fp32'))
for nm in FWD_NAMES:
    got = to_nchw(eng.tensor(nm))
    a, b = res['fp32'][0][nm], res['bf16'][0][nm]
    print('%-14s %12.3e %12.3e %12.3e' % (nm, nrel(got.reshape(a.shape), a), nrel(got.reshape(b.shape), b), nrel(b, a)))
print('losses native', model.last_losses()); print('losses fp32  ', res['fp32'][2]); print('losses emu   ', res['bf16'][2])
for nm in GRAD_NAMES:
    got = to_nchw(eng.tensor(nm))
    a, b = res['fp32'][1][nm], res['bf16'][1][nm]
    print('%-14s %12.3e %12.3e %12.3e' % (nm, nrel(got.reshape(a.shape), a), nrel(got.reshape(b.shape), b), nrel(b, a)))
for k in model.model._adapt_names:
    if k in ZERO_GRAD: continue
    got = model.model._grad_views[k].cpu()
    a, b = res['fp32'][3][k], res['bf16'][3][k]
    print('%-44s %12.3e %12.3e %12.3e' % (k, nrel(got, a), nrel(got, b), nrel(b, a)))
