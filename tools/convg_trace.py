#!/usr/bin/env python
"""[needs the experiments build: python -m tta_depth_completion_b200.build experiments, then run with
PTTA_B200_LIB=tta_depth_completion_b200/lib/libptta_b200_experiments.so -- the product library contains none of these switches]
Per-tile cycle trace of CTA 0 of the general-channel conv (64->64 3x3 s1 @352x1216): where does a tile's time go?
events: issue thread 0 loop top | 1 accumulator free | 2 A tile landed | 3 36 MMAs issued + commits;
epilogue 4 waiting | 5 accumulator full | 6 tile staged in smem | 7 accumulator released; producer 8 loop top | 9 TMA issued"""
import ctypes
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tta_depth_completion_b200 import _lib
from tta_depth_completion_b200.convg import ConvG, FWD

dev = torch.device('cuda:0')
L = _lib.lib()
c, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 352, 1216)
op = ConvG('s1', FWD, torch.randn((c, c, 3, 3), device=dev) * 0.05, c, c)
x = torch.randn((1, h, w, c), device=dev).to(torch.bfloat16)
out = torch.empty_like(x)
for _ in range(3):
    op(x, out=out)
L.ptta_convg_debug_set(64)
op(x, out=out)
buf = (ctypes.c_longlong * (64 * 16))()
_lib.check(L.ptta_convg_debug_read_ts(buf, 64 * 16), 'read_ts')
L.ptta_convg_debug_set(0)
cta = (ctypes.c_ulonglong * 296)()
_lib.check(L.ptta_convg_debug_read_cta(cta, 296), 'read_cta')
spans = [(cta[2 * i], cta[2 * i + 1]) for i in range(148) if cta[2 * i + 1] > 0]
if spans:
    first = min(a for a, b in spans)
    durs = sorted((b - a) / 1e3 for a, b in spans)
    print('CTAs: %d; start spread %.2f us; busy min %.2f median %.2f max %.2f us; last exit at %.2f us after the first entry' % (
        len(spans), (max(a for a, b in spans) - first) / 1e3, durs[0], durs[len(durs) // 2], durs[-1], (max(b for a, b in spans) - first) / 1e3))
    late = sorted(((b - first) / 1e3, i) for i, (a, b) in enumerate(spans))[-5:]
    print('latest exits (us, cta):', [(round(t, 2), i) for t, i in late])
ts = [[buf[t * 16 + k] for k in range(16)] for t in range(64)]
t0 = ts[0][0]
print('kernel entry %d, set-up done %d, first tile top 0, all roles finished %d (cycles relative to the first tile)' % (ts[63][0] - t0, ts[63][1] - t0, ts[63][2] - t0))
if ts[8][0] == 0:          # few tiles per CTA (streamed-weight layers): print everything, no steady-state summary
    print('tile | issue: top accfree Aready(last chunk) issued(last chunk) | chunk 0: wait/got B taps 0, 3, 6 | epi: wait full staged released | prod: top tma')
    for t in range(64):
        r = ts[t]
        if r[0] == 0 and t > 0:
            break
        print('%3d | %7d %7d %7d %7d | %7d %7d  %7d %7d  %7d %7d | %7d %7d %7d %7d | %7d %7d' % (
            (t,) + tuple(v - t0 for v in r[:4]) + tuple(v - t0 for v in r[10:16]) + tuple(v - t0 for v in r[4:8]) + tuple(v - t0 for v in r[8:10])))
    sys.exit(0)
print('tile | issue: top accfree Aready issued | epi: wait full staged released | prod: top tma   (cycles since the first stamp)')
for t in range(2, 24):
    r = ts[t]
    if r[0] == 0:
        break
    print('%4d | %7d %7d %7d %7d | %7d %7d %7d %7d | %7d %7d' % ((t,) + tuple(v - t0 for v in r[:10])))
per = (ts[20][0] - ts[4][0]) / 16.0
print('cycles per tile (issue thread, tiles 4..20): %.0f' % per)
for name, a, b in (('wait accumulator', 0, 1), ('wait A tile', 1, 2), ('issue 36 MMAs', 2, 3)):
    print('  %-18s %.0f' % (name, sum(ts[t][b] - ts[t][a] for t in range(4, 20)) / 16.0))
print('  %-18s %.0f' % ('issued -> next top', sum(ts[t + 1][0] - ts[t][3] for t in range(4, 20)) / 16.0))
print('epilogue: full -> staged %.0f, staged -> released %.0f, released -> next full %.0f' % (
    sum(ts[t][6] - ts[t][5] for t in range(4, 20)) / 16.0, sum(ts[t][7] - ts[t][6] for t in range(4, 20)) / 16.0,
    sum(ts[t + 1][5] - ts[t][7] for t in range(4, 20)) / 16.0))
print('issue commit -> epilogue sees full: %.0f' % (sum(ts[t][5] - ts[t][3] for t in range(4, 20)) / 16.0))
