#!/usr/bin/env python
"""GPU experiment (context for DESIGN.md only, NOT a bench arm): the oracle's PyTorch restatement of the step run in eager
fp32 on the B200 through cuDNN/cuBLAS -- the closest stand-in available on the GPU box for "the reference's existing GPU
path" (the reference itself cannot travel).  It back-propagates only to the adapted tensors, so it does LESS work than the
reference's autograd (225 vs 400 GFLOP): an optimistic estimate of the reference."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import msgchn_oracle as O
dev = torch.device('cuda:0')
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    h, w, dataset, mode, lr, cap = bench.WORKLOADS['kitti']
    sd = {k: v.to(dev) for k, v in bench.make_checkpoint('kitti').items()}
    names = O.adapt_parameter_names(sd, 'meta')
    state = O.AdamState(names, sd)
    frames = [(i.to(dev), s.to(dev)) for i, s in bench.make_frames('kitti', 1, 4, 1)]
    for i in range(3):
        O.tta_step(sd, state, *frames[i % 4], lr=lr, max_input_depth=cap)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for i in range(n):
        O.tta_step(sd, state, *frames[i % 4], lr=lr, max_input_depth=cap)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print('torch %s eager on %s, tf32=%s: %.2f ms/step = %.1f frames/s' % (torch.__version__, torch.cuda.get_device_name(0), tf32, dt * 1e3, 1 / dt), flush=True)
