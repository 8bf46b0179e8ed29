#!/usr/bin/env python
"""GPU experiment: per-row cycle stamps of the tcgen05 conv pipeline (CTA 0): where does each role wait?"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops, _lib
dev = 'cuda'
g = torch.Generator().manual_seed(0)
wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wp = ops.pack_conv_weight(wt, 'conv_fwd'); wi = ops.pack_conv_weight_tc(wp); bias = torch.zeros(32, device=dev)
n, h, w = 1, 352, 1216
xs = [torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16) for _ in range(4)]
extra = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for i in range(3):
    ops.conv3x3_tc(xs[i], wp, bias, relu_in=False, wimage=wi)
_lib.lib().ptta_debug_set(64 | extra)
ops.conv3x3_tc(xs[3], wp, bias, relu_in=False, wimage=wi)
torch.cuda.synchronize()
_lib.lib().ptta_debug_set(0)
buf = (ctypes.c_longlong * (3 * 2048))()
_lib.check(_lib.lib().ptta_debug_read_ts(buf, 3 * 2048))
ts = list(buf)
t0 = min(t for t in ts if t > 0)
names = {0: 'MMA  [start, slot_empty ok, row_full ok, issued+committed]',
         1: 'EPI  [start, slot_full ok, tmem ld+zero+arrive, stored]',
         2: 'TMA  [start, row_free ok, issued]'}
for role in range(3):
    print(names[role])
    prev = None
    for row in range(40):
        v = ts[role * 2048 + row * 8: role * 2048 + row * 8 + 6] if not (role == 2 and row >= 255) else [0] * 6
        if not any(v):
            continue
        rel = [x - t0 if x else -1 for x in v]
        d = [rel[k + 1] - rel[k] if rel[k + 1] >= 0 and rel[k] >= 0 else -1 for k in range(5)]
        print('  row %3d t=%7d  steps %s  (since prev row start %s)' % (row, rel[0], d, rel[0] - prev if prev is not None else '-'))
        prev = rel[0]
k = ts[3 * 2048 - 8: 3 * 2048 - 5]
print('kernel entry -> setup done %d cycles; entry -> teardown %d cycles; first stamp at %d after entry' % (k[1] - k[0], k[2] - k[0], t0 - k[0]))
