import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops
dev='cuda'
g = torch.Generator().manual_seed(0)
wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wp = ops.pack_conv_weight(wt, 'conv_fwd'); bias = torch.zeros(32, device=dev)
for rep in range(3):
    for (h,w) in [(352,1216),(176,608),(88,304),(44,152),(22,76)]:
        x=torch.randn((1,h,w,32),device=dev).to(torch.bfloat16)
        ops.conv3x3_tc(x, wp, bias, relu_in=True)
        ops.conv3x3(x, wp, bias, ops.MODE_S1, ops.PRO_RELU)
torch.cuda.synchronize()
