#!/usr/bin/env python
"""[needs the experiments build: python -m tta_depth_completion_b200.build experiments, then run with
PTTA_B200_LIB=tta_depth_completion_b200/lib/libptta_b200_experiments.so -- the product library contains none of these switches]
Where does the time of the general-channel tcgen05 conv go?  Times the 64->64 and 256->256 stride-1 layers with parts of the
kernel switched off (ptta_convg_debug_set)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tta_depth_completion_b200 import _lib
from tta_depth_completion_b200.convg import ConvG, FWD

dev = torch.device('cuda:0')
L = _lib.lib()


def timeit(op, xs, outs, reps=20):
    """GPU time per launch: the launches are replayed from a CUDA graph (no host cost between them)"""
    for i in range(3):
        op(xs[i % len(xs)], out=outs[i % len(xs)])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            op(xs[i % len(xs)], out=outs[i % len(xs)])
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for c, h, w in ((64, 352, 1216), (128, 176, 608), (256, 88, 304)):
    wt = torch.randn((c, c, 3, 3), device=dev) * 0.05
    op = ConvG('s1', FWD, wt, c, c)
    xs = [torch.randn((1, h, w, c), device=dev).to(torch.bfloat16) for _ in range(6)]
    outs = [torch.empty((1, h, w, c), dtype=torch.bfloat16, device=dev) for _ in range(6)]
    line = '%d->%d @%dx%d:' % (c, c, h, w)
    for mask, name in ((0, 'full'), (1, 'one MMA/item'), (2, 'no epilogue'), (4, 'no fence/store'), (3, 'one MMA + no epilogue'), (8, 'no A loads'), (16, 'no B loads'), (24, 'no loads'), (27, 'nothing')):
        L.ptta_convg_debug_set(mask)
        line += '  %s %.1f us' % (name, timeit(op, xs, outs))
    L.ptta_convg_debug_set(0)
    print(line)
