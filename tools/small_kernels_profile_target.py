#!/usr/bin/env python
"""ncu target: the CUDA-core kernels of the MSG-CHN step at 352x1216 (stems, prediction conv and its adjoints, up2 add, decoder sums)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tta_depth_completion_b200 import _lib
from tta_depth_completion_b200._lib import ptr, c_void_p, check
L = _lib.lib()
dev = 'cuda'; n, h, w = 1, 352, 1216
s = c_void_p(torch.cuda.current_stream().cuda_stream)
img = torch.rand((n, 3, h, w), device=dev)
wt3 = torch.randn((32, 3, 3, 3), device=dev) * 0.2; wt1 = torch.randn((32, 1, 3, 3), device=dev) * 0.2
b = torch.zeros(32, device=dev)
out = torch.empty((n, h, w, 32), dtype=torch.bfloat16, device=dev)
mask = torch.randn((n, h, w, 32), device=dev).to(torch.bfloat16)
hw = h * w
for rep in range(2):
    planes = (ctypes.c_void_p * 3)(img.data_ptr(), img.data_ptr() + 4 * hw, img.data_ptr() + 8 * hw)
    strides = (ctypes.c_longlong * 3)(3 * hw, 3 * hw, 3 * hw)
    sc = (ctypes.c_float * 3)(1 / 255., 1 / 255., 1 / 255.); sh = (ctypes.c_float * 3)(0, 0, 0)
    wh = wt3.cpu().contiguous(); bh = b.cpu().contiguous()
    check(L.ptta_stem_conv_const(planes, strides, sc, sh, 3, ctypes.c_void_p(wh.data_ptr()), ctypes.c_void_p(bh.data_ptr()), None, ptr(out), 1, n, h, w, s), 'stem3')
    g1 = torch.rand((n, 1, h, w), device=dev)
    planes1 = (ctypes.c_void_p * 3)(g1.data_ptr(), g1.data_ptr(), g1.data_ptr()); strides1 = (ctypes.c_longlong * 3)(hw, hw, hw)
    check(L.ptta_stem_conv(planes1, strides1, sc, sh, 1, ptr(wt1), None, ptr(mask), ptr(out), n, h, w, s), 'stem1 mask (head dgrad)')
    w9 = torch.randn((9, 32), device=dev) * 0.1
    o1 = torch.empty((n, h, w), device=dev)
    w9h = w9.cpu().contiguous()
    check(L.ptta_head_conv_const(ptr(mask), ctypes.c_void_p(w9h.data_ptr()), 0.1, None, ptr(o1), n, h, w, 1, 0, s), 'head_conv')
    half = torch.randn((n, h // 2, w // 2, 32), device=dev).to(torch.bfloat16)
    check(L.ptta_add_up2_c32(ptr(mask), ptr(half), ptr(out), n, h // 2, w // 2, s), 'add_up2')
    glo = torch.empty((n, h // 2, w // 2, 32), dtype=torch.bfloat16, device=dev)
    check(L.ptta_up2_c32_adjoint(ptr(mask), ptr(glo), n, h // 2, w // 2, 0, s), 'up2_adj')
torch.cuda.synchronize()
