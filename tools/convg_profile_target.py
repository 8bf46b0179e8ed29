#!/usr/bin/env python
"""ncu target: a few launches of the general-channel tcgen05 conv at the dominant NLSPN shape (64->64 3x3 s1 @352x1216) and at
256->256 @88x304, on fresh inputs.  `ncu --set full -k regex:convg_kernel -c 4 python tools/convg_profile_target.py`"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tta_depth_completion_b200.convg import ConvG, FWD  # noqa: E402

dev = torch.device('cuda:0')
for c, h, w in ((64, 352, 1216), (256, 88, 304)):
    wt = torch.randn((c, c, 3, 3), device=dev) * 0.05
    op = ConvG('s1', FWD, wt, c, c)
    for i in range(2):
        x = torch.randn((1, h, w, c), device=dev).to(torch.bfloat16)
        op(x)
torch.cuda.synchronize()
