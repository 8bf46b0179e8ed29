#!/usr/bin/env python
"""Times the general-channel tcgen05 conv (csrc/conv_gen.cuh) on the NLSPN layer shapes at KITTI size (1x352x1216 input),
CUDA events over a ring of inputs larger than L2; prints TFLOP/s and the algorithmic GB/s per layer."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tta_depth_completion_b200.convg import ConvG, FWD, DGRAD  # noqa: E402

dev = torch.device('cuda:0')
CASES = [
    # name, kind, role, cin (stored, tuple for concat), cout, h, w (layer input)
    ('layer1 64->64 s1 @352x1216', 's1', FWD, 64, 64, 352, 1216),
    ('layer1 dgrad', 's1', DGRAD, 64, 64, 352, 1216),
    ('layer2.0 64->128 s2', 's2', FWD, 64, 128, 352, 1216),
    ('layer2 128->128 @176x608', 's1', FWD, 128, 128, 176, 608),
    ('layer3 256->256 @88x304', 's1', FWD, 256, 256, 88, 304),
    ('layer4 512->512 @44x152', 's1', FWD, 512, 512, 44, 152),
    ('conv6 512->512 s2 @44x152', 's2', FWD, 512, 512, 44, 152),
    ('dec5 convT 512->256 @22x76', 't2', FWD, 512, 256, 22, 76),
    ('dec4 convT 768->128 @44x152', 't2', FWD, (256, 512), 128, 44, 152),
    ('dec3 convT 384->64 @88x304', 't2', FWD, (128, 256), 64, 88, 304),
    ('dec2 convT 192->64 @176x608', 't2', FWD, (64, 128), 64, 176, 608),
    ('id_dec1 128->64 @352x1216', 's1', FWD, (64, 64), 64, 352, 1216),
    ('id_dec1 dgrad 64->128', 's1', DGRAD, 128, 64, 352, 1216),
    ('dec2 dgrad', 't2', DGRAD, 192, 64, 176, 608),
    ('layer2.0 dgrad + shortcut', 's2', DGRAD, 64, 128, 352, 1216),
]
RING = 6


def main():
    print('%-34s %9s %9s %9s' % ('layer', 'us', 'TFLOP/s', 'GB/s'))
    for name, kind, role, cin, cout, h, w in CASES:
        c0, c1 = (cin, 0) if isinstance(cin, int) else cin
        c = c0 + c1
        shape = (c, cout, 3, 3) if kind == 't2' else (cout, c, 3, 3)
        wt = torch.randn(shape, device=dev) * 0.05
        ws = torch.randn((cout, c, 1, 1), device=dev) * 0.05 if 'shortcut' in name else None
        op = ConvG(kind, role, wt, cin if c1 else c0, cout, weight_short=ws)
        oh, ow = (h, w) if kind == 's1' else ((2 * h, 2 * w) if kind == 't2' else (h // 2, w // 2))
        if role == FWD:
            xs = [(torch.randn((1, h, w, c0), device=dev).to(torch.bfloat16),
                   torch.randn((1, h, w, c1), device=dev).to(torch.bfloat16) if c1 else None) for _ in range(RING)]
            run = lambda i: op(xs[i][0], xs[i][1], out=outs[i])
            outs = [torch.empty((1, oh, ow, cout), dtype=torch.bfloat16, device=dev) for _ in range(RING)]
            in_elems, out_elems = h * w * c, oh * ow * cout
        else:
            xs = [(torch.randn((1, oh, ow, cout), device=dev).to(torch.bfloat16),
                   torch.randn((1, oh, ow, cout), device=dev).to(torch.bfloat16) if ws is not None else None) for _ in range(RING)]
            outs = [torch.empty((1, h, w, c), dtype=torch.bfloat16, device=dev) for _ in range(RING)]
            run = lambda i: op(xs[i][0], xs[i][1], out=outs[i], hw=(h, w))
            in_elems, out_elems = oh * ow * cout * (2 if ws is not None else 1), h * w * c
        taps = 9
        flops = 2.0 * taps * c * cout * (h * w if kind in ('s1', 't2') else oh * ow)
        if ws is not None:
            flops += 2.0 * c * cout * oh * ow
        for i in range(3):
            run(i % RING)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 24
        e0.record()
        for i in range(reps):
            run(i % RING)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print('%-34s %9.1f %9.1f %9.1f' % (name, us, flops / us / 1e6, (in_elems + out_elems) * 2 / us / 1e3))


if __name__ == '__main__':
    main()
