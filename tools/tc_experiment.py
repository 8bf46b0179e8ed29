#!/usr/bin/env python
"""GPU experiment: tcgen05 conv variants vs the mma.sync kernel (correctness + timing). Run each variant in its own
process under `timeout` so that a deadlocked kernel cannot hang the box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops

def main(variant, shapes):
    dev = 'cuda'
    g = torch.Generator().manual_seed(0)
    wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
    wp = ops.pack_conv_weight(wt, 'conv_fwd'); wi = ops.pack_conv_weight_tc(wp)
    bias = (torch.randn(32, generator=g) * 0.1).to(dev)
    for (n, h, w) in shapes:
        x = torch.randn((n, h, w, 32), generator=g).to(dev).to(torch.bfloat16)
        m = torch.randn((n, h, w, 32), generator=g).to(dev).to(torch.bfloat16)
        a = torch.randn((n, h, w, 32), generator=g).to(dev).to(torch.bfloat16)
        for relu_in, use_ma in ((True, False), (False, True)):
            want = ops.conv3x3(x, wp, bias, ops.MODE_S1, ops.PRO_RELU if relu_in else ops.PRO_NONE,
                               mask=m if use_ma else None, mask_mode=ops.MASK_RELU if use_ma else ops.MASK_NONE, add=a if use_ma else None)
            xin = torch.relu(x) if relu_in else x.clone()     # producers store ReLU(x): the tcgen05 kernel has no ReLU-on-load
            got = ops.conv3x3_tc(xin, wp, bias, relu_in=False, mask=m if use_ma else None, add=a if use_ma else None, variant=variant)
            torch.cuda.synchronize()
            err = (got.float() - want.float()).abs()
            tol = 2.0 ** -7 * want.float().abs().clamp_min(float(want.float().pow(2).mean().sqrt()))
            bad = int((err > tol).sum())
            print('variant %d shape %s relu_in=%d mask/add=%d: max err %.4g, bad %d / %d' % (variant, (n, h, w), relu_in, use_ma, float(err.max()), bad, err.numel()), flush=True)
    # timing at full resolution
    n, h, w = 1, 352, 1216
    xs = [torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16) for _ in range(8)]
    mk = torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16)
    for name, fn in (('tc', lambda x: ops.conv3x3_tc(x, wp, bias, relu_in=False, variant=variant, wimage=wi)),
                     ('tc_mask_add', lambda x: ops.conv3x3_tc(x, wp, bias, relu_in=False, mask=mk, add=mk, wimage=wi)),
                     ('mma', lambda x: ops.conv3x3(x, wp, bias, ops.MODE_S1, ops.PRO_RELU))):
        for i in range(5):
            fn(xs[i % 8])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            fn(xs[i % 8])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 40 * 1e3
        print('variant %d %-10s %.1f us  %.0f TFLOP/s  %.0f GB/s' % (variant, name, us, 7.889 / us * 1e3, 54.8e6 / us / 1e3), flush=True)

if __name__ == '__main__':
    v = int(sys.argv[1])
    main(v, [(1, 16, 128), (1, 24, 300), (2, 19, 38), (1, 88, 304), (1, 5, 1216), (3, 64, 514)])
