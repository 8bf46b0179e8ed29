#!/usr/bin/env python
"""ncu target: a few launches of the tcgen05 conv at 352x1216 over rotating inputs (the roofline kernel of bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops
dev = 'cuda'
g = torch.Generator().manual_seed(0)
wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wi = ops.pack_conv_weight_tc(ops.pack_conv_weight(wt, 'conv_fwd'))
bias = torch.zeros(32, device=dev)
xs = [torch.relu(torch.randn((1, 352, 1216, 32), device=dev)).to(torch.bfloat16) for _ in range(8)]
mk = torch.randn((1, 352, 1216, 32), device=dev).to(torch.bfloat16)
for i in range(6):
    ops.conv3x3_tc(xs[i % 8], None, bias, wimage=wi)
for i in range(2):
    ops.conv3x3_tc(xs[i % 8], None, None, mask=mk, add=mk, wimage=wi)
torch.cuda.synchronize()
