#!/usr/bin/env python
"""CPU study (no kernel involved): how far does rounding the stored activations / weights / gradient maps to bf16 (or fp16) move the
gradients and the adapted weights of the MSG-CHN TTA step, on a given checkpoint?  Uses the oracle's own Precision hook.
    python tools/precision_study.py [ckpt] [h] [w] [steps]      ckpt: fitted name (kitti_2layers_a ...) or integer seed"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import msgchn_oracle as O

ckpt = sys.argv[1] if len(sys.argv) > 1 else 'kitti_2layers_a'
ckpt = int(ckpt) if ckpt.lstrip('-').isdigit() else ckpt
h = int(sys.argv[2]) if len(sys.argv) > 2 else 64
w = int(sys.argv[3]) if len(sys.argv) > 3 else 128
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mode = 'meta_selfsup_seq_1layer_ema' if isinstance(ckpt, str) and 'void' in ckpt else 'meta_selfsup_seq_2layers_ema'
dataset, cap, lr = ('void', 8.0, 3e-3) if 'void' in str(ckpt) else ('kitti', 80.0, 1e-4)
torch.set_num_threads(os.cpu_count())


class P16(O.Precision):
    def act(self, x):
        return x.to(torch.float16).to(torch.float32) if self.emulate == 'fp16' else super().act(x)

    def wgt(self, x):
        return x.to(torch.float16).to(torch.float32) if self.emulate == 'fp16' else super().wgt(x)


def nrel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def run(pr):
    sd = O.get_checkpoint(ckpt, mode)
    sd0 = {k: v.clone() for k, v in sd.items()}
    names = O.adapt_parameter_names(sd)
    st = O.AdamState(names, sd)
    outs = []
    for t in range(steps):
        image, sparse, dense = O.synthetic_frame(11, t, 1, h, w, dataset)
        r = O.tta_step(sd, st, image, sparse, lr=lr, max_input_depth=cap, pr=pr, return_grads=True)
        outs.append(r)
    return sd0, sd, outs, names


sd0, ref_sd, ref, names = run(O.FP32)
print('checkpoint %s  %dx%d  %d steps  losses step0: loss %.5f sd %.5f sm %.5f cos %.5f' % (
    ckpt, h, w, steps, ref[0]['loss'], ref[0]['loss_sparse_depth'], ref[0]['loss_smooth'], ref[0]['loss_cos']))
for tag in ('bf16', 'fp16'):
    _, sd, outs, _ = run(P16(tag))
    print('--- %s emulation vs fp32' % tag)
    for t in range(steps):
        print('  step %d: loss rel %.2e  cos rel %.2e  output nrel %.2e' % (
            t, abs(outs[t]['loss'] - ref[t]['loss']) / abs(ref[t]['loss']), abs(outs[t]['loss_cos'] - ref[t]['loss_cos']) / abs(ref[t]['loss_cos']),
            nrel(outs[t]['output_depth'], ref[t]['output_depth'])))
    for k in names:
        print('  %-44s grad nrel step0 %.3e   weight nrel %.3e   update/|w| %.3e' % (
            k, nrel(outs[0]['grads'][k], ref[0]['grads'][k]), nrel(sd[k], ref_sd[k]), nrel(sd0[k], ref_sd[k])))
