#!/usr/bin/env python
"""GPU experiment: stride-2 tcgen05 conv vs the mma.sync kernel (bit-exactness + graph-replayed timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops
dev = 'cuda'
g = torch.Generator().manual_seed(0)
wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wp = ops.pack_conv_weight(wt, 'conv_fwd')
bias = (torch.randn(32, generator=g) * 0.1).to(dev)
for (n, h, w) in [(1, 16, 256), (1, 32, 48), (2, 18, 38), (1, 88, 304), (1, 176, 608), (3, 64, 516), (1, 352, 1216)]:
    x = torch.randn((n, h, w, 32), generator=g).to(dev).to(torch.bfloat16)
    m = torch.randn((n, h // 2, w // 2, 32), generator=g).to(dev).to(torch.bfloat16)
    a = torch.randn((n, h // 2, w // 2, 32), generator=g).to(dev).to(torch.bfloat16)
    want = ops.conv3x3(x, wp, bias, ops.MODE_S2, ops.PRO_RELU)
    got, got2 = ops.conv3x3_tc_s2(torch.relu(x), wp, bias, want_relu_copy=True)
    torch.cuda.synchronize()
    bad = int((got != want).sum()); bad2 = int((got2 != torch.relu(want)).sum())
    want_ma = ops.conv3x3(x, wp, None, ops.MODE_S2, ops.PRO_NONE, mask=m, mask_mode=ops.MASK_RELU, add=a)
    got_ma = ops.conv3x3_tc_s2(x, wp, None, mask=m, add=a)
    bad3 = int((got_ma != want_ma).sum())
    print('shape %s: fwd mismatches %d / %d (max err %.3g), relu copy %d, mask+add %d' % ((n, h, w), bad, got.numel(),
          float((got.float() - want.float()).abs().max()), bad2, bad3), flush=True)


def graph_time(fn, xs, iters=40):
    for i in range(3):
        fn(xs[i % len(xs)])
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for i in range(iters):
                fn(xs[i % len(xs)])
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); gr.replay(); e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for (n, h, w) in [(1, 352, 1216), (1, 176, 608), (1, 88, 304)]:
    xs = [torch.relu(torch.randn((n, h, w, 32), device=dev)).to(torch.bfloat16) for _ in range(8)]
    mk = torch.randn((n, h // 2, w // 2, 32), device=dev).to(torch.bfloat16)
    t_tc = graph_time(lambda x: ops.conv3x3_tc_s2(x, wp, bias), xs)
    t_tcm = graph_time(lambda x: ops.conv3x3_tc_s2(x, wp, None, mask=mk), xs)
    t_mma = graph_time(lambda x: ops.conv3x3(x, wp, bias, ops.MODE_S2, ops.PRO_RELU), xs)
    t_mmam = graph_time(lambda x: ops.conv3x3(x, wp, None, ops.MODE_S2, ops.PRO_NONE, mask=mk, mask_mode=ops.MASK_RELU), xs)
    gb = n * h * w * 64 * 1.25 / 1e3
    print('%s  tc %.1f us (%.0f GB/s, incl. the per-call weight-image kernel) | tc+mask %.1f | mma %.1f | mma+mask %.1f' % ((n, h, w), t_tc, gb / t_tc, t_tcm, t_mma, t_mmam), flush=True)
