#!/usr/bin/env python
"""GPU experiment: the CUDA-core kernels of the MSG-CHN step at 352x1216, each timed from a CUDA graph of back-to-back launches over
rotating buffers (> L2), with the algorithmic bytes they move and the resulting fraction of the measured HBM peak."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import _lib
from tta_depth_completion_b200._lib import ptr, c_void_p, check
L = _lib.lib()
dev = 'cuda'; n, h, w = 1, 352, 1216; hw = h * w
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    PEAK = 6456.2
R = 6
imgs = [torch.rand((n, 3, h, w), device=dev) for _ in range(R)]
maps = [torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16) for _ in range(R)]
outs = [torch.empty((n, h, w, 32), dtype=torch.bfloat16, device=dev) for _ in range(R)]
halves = [torch.randn((n, h // 2, w // 2, 32), device=dev).to(torch.bfloat16) for _ in range(R)]
o1 = [torch.empty((n, h, w), device=dev) for _ in range(R)]
wt3 = torch.randn((32, 3, 3, 3), device=dev) * 0.2; wt2 = torch.randn((32, 2, 3, 3), device=dev) * 0.2; wt1 = torch.randn((32, 1, 3, 3), device=dev) * 0.2
b = torch.zeros(32, device=dev); w9 = torch.randn((9, 32), device=dev) * 0.1
sc = (ctypes.c_float * 3)(1 / 255., 1 / 255., 1 / 255.); sh = (ctypes.c_float * 3)(0, 0, 0)


def stem(cin, wt, mask):
    def f(i, s):
        im = imgs[i % R]
        planes = (ctypes.c_void_p * 3)(im.data_ptr(), im.data_ptr() + 4 * hw, im.data_ptr() + 8 * hw)
        strides = (ctypes.c_longlong * 3)(3 * hw, 3 * hw, 3 * hw)
        check(L.ptta_stem_conv(planes, strides, sc, sh, cin, ptr(wt), ptr(b), ptr(maps[(i + 1) % R]) if mask else None, ptr(outs[i % R]), n, h, w, s), 'stem')
    return f


def stemc(cin, wt, mask):
    wh = wt.cpu().contiguous(); bh = b.cpu().contiguous()
    def f(i, s):
        im = imgs[i % R]
        planes = (ctypes.c_void_p * 3)(im.data_ptr(), im.data_ptr() + 4 * hw, im.data_ptr() + 8 * hw)
        strides = (ctypes.c_longlong * 3)(3 * hw, 3 * hw, 3 * hw)
        check(L.ptta_stem_conv_const(planes, strides, sc, sh, cin, ctypes.c_void_p(wh.data_ptr()), ctypes.c_void_p(bh.data_ptr()),
                                     ptr(maps[(i + 1) % R]) if mask else None, ptr(outs[i % R]), 0, n, h, w, s), 'stemc')
    return f


w9h = w9.cpu().contiguous()
img_tc = {c: torch.empty(2048, dtype=torch.bfloat16, device=dev) for c in (1, 2, 3)}
head_img = torch.empty(9 * 16 * 32, dtype=torch.bfloat16, device=dev)
check(L.ptta_pack_head_weight_tc(ptr(w9), ptr(head_img), None), 'pack_head_tc')


def stemtc(cin, wt, mask):
    first = [True]
    def f(i, s):
        im = imgs[i % R]
        planes = (ctypes.c_void_p * 3)(im.data_ptr(), im.data_ptr() + 4 * hw, im.data_ptr() + 8 * hw)
        strides = (ctypes.c_longlong * 3)(3 * hw, 3 * hw, 3 * hw)
        check(L.ptta_stem_conv_tc(planes, strides, sc, sh, cin, ptr(wt) if first[0] else None, ptr(b), ptr(maps[(i + 1) % R]) if mask else None, ptr(outs[i % R]),
                                  ptr(img_tc[cin]), 0 if mask else 1, n, h, w, s), 'stem_tc')
        first[0] = False
    return f


CASES = [
    ('stem_tc 3->32 (tcgen05)', stemtc(3, wt3, False), 12 * hw + 64 * hw),
    ('stem_tc 2->32 (tcgen05)', stemtc(2, wt2, False), 8 * hw + 64 * hw),
    ('head_dgrad stem_tc (1->32 + ReLU mask)', stemtc(1, wt1, True), 4 * hw + 128 * hw),
    ('head_conv_tc 32->1 (tcgen05)', lambda i, s: check(L.ptta_head_conv_tc(ptr(maps[i % R]), ptr(head_img), 0.1, None, ptr(o1[i % R]), n, h, w, s), 'head_tc'), 64 * hw + 4 * hw),
    ('stem_conv_const 3->32', stemc(3, wt3, False), 12 * hw + 64 * hw),
    ('stem_conv_const 2->32', stemc(2, wt2, False), 8 * hw + 64 * hw),
    ('head_dgrad const (1->32 + ReLU mask)', stemc(1, wt1, True), 4 * hw + 128 * hw),
    ('head_conv_const 32->1', lambda i, s: check(L.ptta_head_conv_const(ptr(maps[i % R]), ctypes.c_void_p(w9h.data_ptr()), 0.1, None, ptr(o1[i % R]), n, h, w, 1, 0, s), 'headc'), 64 * hw + 4 * hw),
    ('stem_conv 3->32', stem(3, wt3, False), 12 * hw + 64 * hw),
    ('stem_conv 2->32', stem(2, wt2, False), 8 * hw + 64 * hw),
    ('head_dgrad (stem 1->32 + ReLU mask)', stem(1, wt1, True), 4 * hw + 128 * hw),
    ('head_conv 32->1', lambda i, s: check(L.ptta_head_conv(ptr(maps[i % R]), ptr(w9), 0.1, None, ptr(o1[i % R]), n, h, w, 1, 0, s), 'head'), 64 * hw + 4 * hw),
    ('add_up2_c32', lambda i, s: check(L.ptta_add_up2_c32(ptr(maps[i % R]), ptr(halves[i % R]), ptr(outs[i % R]), n, h // 2, w // 2, s), 'add_up2'), 128 * hw + 16 * hw),
    ('up2_c32_adjoint', lambda i, s: check(L.ptta_up2_c32_adjoint(ptr(maps[i % R]), ptr(halves[i % R]), n, h // 2, w // 2, 0, s), 'up2_adj'), 64 * hw + 16 * hw),
]
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    s = c_void_p(st.cuda_stream)
    for name, fn, nbytes in CASES:
        for i in range(3):
            fn(i, s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            s2 = c_void_p(torch.cuda.current_stream().cuda_stream)
            for i in range(24):
                fn(i, s2)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 24
        print('%-38s %6.1f us  %6.1f MB  %5.0f GB/s  %.2f of HBM peak' % (name, us, nbytes / 1e6, nbytes / us / 1e3, nbytes / us / 1e3 / PEAK), flush=True)
