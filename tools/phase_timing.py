#!/usr/bin/env python
"""GPU experiment: where does a TTA step spend its time?  Eager launches, CUDA events between the phases of the step
(outlier removal + forward | loss | backward | Adam + repack), 352x1216, averaged over 50 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import msgchn_oracle as O
from tta_depth_completion_b200 import ExternalModel_Adapt, ops
dev = torch.device('cuda:0')
h, w, dataset, mode, lr, cap = bench.WORKLOADS['kitti']
model = ExternalModel_Adapt('msg_chn', 0.0, 100.0, max_input_depth=cap, device=dev)
model._prepare_head(mode)
model.load_state_dict(bench.make_checkpoint('kitti'))
model.set_image_normalization((1 / 255.0,) * 3, (0.0,) * 3)
model.train()
frames = [(i.to(dev), s.to(dev)) for i, s in bench.make_frames('kitti', 1, 4, 1)]
st = torch.cuda.Stream(dev)
with torch.cuda.stream(st):
    for i in range(3):
        model.tta_step(frames[i % 4][0], frames[i % 4][1], lr, 1.0, 1.0, 0.1)
    eng = model._last_engine
    torch.cuda.synchronize()
    names = ['outlier', 'forward', 'loss', 'backward', 'adam+repack']
    tot = [0.0] * len(names)
    steps = 50
    scale, shift = model.model.img_scale, model.model.img_shift
    for it in range(steps):
        img, sp = frames[it % 4]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record(st)
        d, v = ops.outlier_removal(sp)
        ev[1].record(st)
        eng.forward(img, d, cap, True, scale, shift)
        ev[2].record(st)
        eng.loss(img, d, v, cap, 1.0, 1.0, 0.1)
        ev[3].record(st)
        eng.backward(1.0)
        ev[4].record(st)
        eng.adam_step()
        ev[5].record(st)
        torch.cuda.synchronize()
        for k in range(len(names)):
            tot[k] += ev[k].elapsed_time(ev[k + 1])
    for k, nm in enumerate(names):
        print('%-12s %8.1f us' % (nm, 1e3 * tot[k] / steps))
    print('%-12s %8.1f us' % ('sum', 1e3 * sum(tot) / steps))
