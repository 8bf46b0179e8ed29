import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops, _lib
dev='cuda'
g = torch.Generator().manual_seed(0)
wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(dev)
wp = ops.pack_conv_weight(wt, 'conv_fwd'); bias = torch.zeros(32, device=dev)
def timeit(fn, xs, iters=40):
    for i in range(5): fn(xs[i % len(xs)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(xs[i % len(xs)])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for shape in [(1,352,1216),(4,352,1216)]:
    n,h,w=shape
    xs=[torch.randn((n,h,w,32),device=dev).to(torch.bfloat16) for _ in range(max(2, min(8, int(300e6/(n*h*w*64)))))]
    res=[]
    for dbg in (0,1|2|4,1|2|4|8,1|2|4|16,1|2|4|8|16,8,16):
        _lib.lib().ptta_debug_set(dbg)
        us=timeit(lambda x: ops.conv3x3_tc(x, wp, bias, relu_in=False), xs)
        res.append('dbg%d %.1f'%(dbg,us))
    _lib.lib().ptta_debug_set(0)
    us=timeit(lambda x: ops.conv3x3_tc(x, wp, bias, relu_in=True, variant=0), xs); res.append('relu %.1f'%us)
    us=timeit(lambda x: ops.conv3x3(x, wp, bias, ops.MODE_S1, ops.PRO_RELU), xs); res.append('mma %.1f'%us)
    print(shape, ' | '.join(res), flush=True)
