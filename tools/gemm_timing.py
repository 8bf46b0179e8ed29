import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import ops
dev='cuda'
m,n,k=26752,512,512
As=[torch.randn((m,k),device=dev).to(torch.bfloat16) for _ in range(6)]
B=(torch.randn((n,k),device=dev)/k**0.5).to(torch.bfloat16); bias=torch.randn(n,device=dev)
def timeit(fn, iters=30):
    for i in range(5): fn(As[i%6])
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(As[i%6])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters*1e3
ref=(As[0].float()@B.float().t()+bias).to(torch.bfloat16)
got=ops.gemm_bf16_tc(As[0],B,bias)
print('max err vs torch', float((got.float()-ref.float()).abs().max()), 'max', float(ref.float().abs().max()))
for name,fn in (('tc',lambda a: ops.gemm_bf16_tc(a,B,bias)),('mma',lambda a: ops.gemm_bf16(a,B,bias)),('torch',lambda a: torch.nn.functional.linear(a,B,bias.to(torch.bfloat16)))):
    us=timeit(fn); print('%-6s %.1f us  %.0f TFLOP/s'%(name,us,2*m*n*k/us/1e6))
# K = 32 (proj.0 of the heads: 32 -> 512): one zero-filled K block on the tcgen05 kernel vs the mma.sync kernel
A32=[torch.randn((m,32),device=dev).to(torch.bfloat16) for _ in range(6)]
B32=(torch.randn((n,32),device=dev)/32**0.5).to(torch.bfloat16)
for name,fn in (('tc k32',lambda i: ops.gemm_bf16_tc(A32[i],B32,bias)),('mma k32',lambda i: ops.gemm_bf16(A32[i],B32,bias))):
    for i in range(5): fn(i%6)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): fn(i%6)
    e1.record(); torch.cuda.synchronize()
    print('%-8s %.1f us (writes %.1f MB)'%(name, e0.elapsed_time(e1)/30*1e3, m*n*2/1e6))
