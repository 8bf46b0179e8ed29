#!/usr/bin/env python
"""GPU experiment: tcgen05 conv vs mma.sync conv, timed from a CUDA graph of 40 back-to-back launches (no host launch cost)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops, _lib
dev = 'cuda'
g = torch.Generator().manual_seed(0)
wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wp = ops.pack_conv_weight(wt, 'conv_fwd'); wi = ops.pack_conv_weight_tc(wp); bias = torch.zeros(32, device=dev)


def timeit(fn, xs, iters=40, graph=True):
    for i in range(3):
        fn(xs[i % len(xs)])
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        if graph:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for i in range(iters):
                    fn(xs[i % len(xs)])
            run = gr.replay
        else:
            def run():
                for i in range(iters):
                    fn(xs[i % len(xs)])
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        run()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for shape in [(1, 352, 1216), (4, 352, 1216), (1, 176, 608), (1, 88, 304)]:
    n, h, w = shape
    cnt = max(2, min(64, int(300e6 / (n * h * w * 64))))
    xs = [torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16) for _ in range(cnt)]
    mk = torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16)
    res = []
    res.append('tc %.1f' % timeit(lambda x: ops.conv3x3_tc(x, wp, bias, wimage=wi), xs))
    res.append('warm %.1f' % timeit(lambda x: ops.conv3x3_tc(x, wp, bias, wimage=wi), xs[:1]))
    res.append('out2+add2 %.1f' % timeit(lambda x: ops.conv3x3_tc_ex(x, wp, bias, add2=mk, wimage=wi), xs))
    res.append('mask %.1f' % timeit(lambda x: ops.conv3x3_tc(x, wp, bias, mask=mk, wimage=wi), xs))
    res.append('mask+add %.1f' % timeit(lambda x: ops.conv3x3_tc(x, wp, bias, mask=mk, add=mk, wimage=wi), xs))
    res.append('mma %.1f' % timeit(lambda x: ops.conv3x3(x, wp, bias, ops.MODE_S1, ops.PRO_RELU), xs))
    res.append('mma mask+add %.1f' % timeit(lambda x: ops.conv3x3(x, wp, bias, ops.MODE_S1, ops.PRO_NONE, mask=mk, mask_mode=ops.MASK_RELU, add=mk), xs))
    gb = 2 * n * h * w * 64 / 1e3
    print(shape, ' | '.join(res), '| tc %.0f GB/s' % (gb / float(res[0].split()[1])), flush=True)

# transposed stride-2: tcgen05 vs mma.sync (input sizes; output is 2x)
wtT = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wpT = ops.pack_conv_weight(wtT, 'convT_fwd')
for shape in [(1, 176, 608), (1, 88, 304), (1, 44, 152)]:
    n, h, w = shape
    xs = [torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16) for _ in range(8)]
    mk = torch.randn((n, 2 * h, 2 * w, 32), device=dev).to(torch.bfloat16)
    res = ['t2 tc %.1f' % timeit(lambda x: ops.conv3x3_tc_t2(x, wpT, bias), xs),
           't2 tc mask+add %.1f' % timeit(lambda x: ops.conv3x3_tc_t2(x, wpT, None, mask=mk, add=mk), xs),
           't2 mma %.1f' % timeit(lambda x: ops.conv3x3(x, wpT, bias, ops.MODE_T2, ops.PRO_RELU), xs),
           't2 mma mask+add %.1f' % timeit(lambda x: ops.conv3x3(x, wpT, None, ops.MODE_T2, ops.PRO_NONE, mask=mk, mask_mode=ops.MASK_RELU, add=mk), xs)]
    print(shape, ' | '.join(res), flush=True)
