#!/usr/bin/env python
"""GPU experiment: NLSPN propagation (18 steps) at 352x1216 -- fused B200 kernels vs the reference's DCN CUDA kernels
(oracle/_ref/dcn_ref.so, when present), timed from CUDA graphs.  Prints us, GB/s of algorithmic traffic (116 B per pixel per step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nlspn_prop_oracle as P
from tta_depth_completion_b200 import nlspn_prop as NP
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
dev = 'cuda'
n, h, w, T = 1, 352, 1216, 18
feat_init, sparse, offset_aff, conf = (t.to(dev) for t in P.synthetic_prop_inputs(1, n, h, w, offset_std=1.5))
offset_aff[:, 16:] *= 0.3


S = torch.cuda.Stream()
torch.cuda.set_stream(S)          # everything (including autograd's backward) runs on one non-default, capturable stream


def graph_time(fn, iters=10):
    fn(); torch.cuda.synchronize()
    s = S
    if True:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


offset, aff = NP.offset_affinity(offset_aff, conf, 4.0, True)
px = n * h * w
us = graph_time(lambda: NP.offset_affinity(offset_aff, conf, 4.0, True))
print('offset_affinity fwd        %8.1f us   %6.0f GB/s (212 B/px)' % (us, px * 212 / us / 1e3))
oa_r, cf_r = offset_aff.clone().requires_grad_(True), conf.clone().requires_grad_(True)
o_r, a_r = NP.offset_affinity(oa_r, cf_r, 4.0, True)
g_o, g_a = torch.randn_like(o_r), torch.randn_like(a_r)
us = graph_time(lambda: torch.autograd.grad((o_r, a_r), (oa_r, cf_r), (g_o, g_a), retain_graph=True))
print('offset_affinity bwd        %8.1f us   %6.0f GB/s (312 B/px + the confidence scatter)' % (us, px * 312 / us / 1e3))
us = graph_time(lambda: NP.propagate(feat_init, offset, aff, sparse, T))
print('propagate fwd (18 steps)   %8.1f us   %6.0f GB/s (116 B/px/step)  %.1f us/step' % (us, px * 116 * T / us / 1e3, us / T))
fi, of, af = (t.clone().requires_grad_(True) for t in (feat_init, offset, aff))
y = NP.propagate(fi, of, af, sparse, T)
go = torch.randn_like(y)
us = graph_time(lambda: torch.autograd.grad(y, (fi, of, af), go, retain_graph=True))
print('propagate bwd (18 steps)   %8.1f us   %6.0f GB/s (336 B/px/step)  %.1f us/step' % (us, px * 336 * T / us / 1e3, us / T))
try:
    from test_nlspn_gpu import dcn_ref
    ref = dcn_ref()
except Exception as e:
    ref = None
    print('reference kernels unavailable:', e)
if ref is not None:
    ones, zero = torch.ones((1, 1, 3, 3), device=dev), torch.zeros(1, device=dev)
    mask_fix = (sparse > 0).float()

    def ref_loop():
        f = feat_init
        for _ in range(T):
            f = (1.0 - mask_fix) * f + mask_fix * sparse
            f = ref.modulated_deform_conv_forward(f, ones, zero, offset, aff, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
        return f
    yr = ref_loop()
    print('fused vs reference kernels: rel err %.2e' % float((y.detach() - yr).norm() / yr.norm()))
    us = graph_time(ref_loop, iters=3)
    print('reference DCN fwd x18      %8.1f us   (%.1f us/step)' % (us, us / T))
    go1 = torch.randn_like(yr)
    fb = (1.0 - mask_fix) * feat_init + mask_fix * sparse
    us = graph_time(lambda: ref.modulated_deform_conv_backward(fb, ones, zero, offset, aff, go1, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64), iters=3)
    print('reference DCN bwd x1       %8.1f us   (x18 = %.1f us)' % (us, us * T))
