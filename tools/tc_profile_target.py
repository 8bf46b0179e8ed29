import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops
dev='cuda'
g = torch.Generator().manual_seed(0)
wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wp = ops.pack_conv_weight(wt, 'conv_fwd'); bias = torch.zeros(32, device=dev)
n,h,w = 1,352,1216
xs=[torch.randn((n,h,w,32),device=dev).to(torch.bfloat16) for _ in range(8)]
which = sys.argv[1] if len(sys.argv) > 1 else 'tc'
for i in range(6):
    if which == 'tc': ops.conv3x3_tc(xs[i % 8], wp, bias, relu_in=False)
    else: ops.conv3x3(xs[i % 8], wp, bias, ops.MODE_S1, ops.PRO_RELU)
torch.cuda.synchronize()
