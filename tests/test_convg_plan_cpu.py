"""Host logic of the general-channel convolution family on the CPU tier: the K-item plan `convg_make_plan` builds (taps, parity
views, parity classes of the transposed conv, skip-concat sources, folded shortcut, weight-chunk order) is read back through
`ptta_convg_plan_describe` and the implicit GEMM it describes is replayed in PyTorch fp64 -- it must equal F.conv2d /
F.conv_transpose2d and the data gradients autograd computes.  No kernel runs here (the GPU tier checks the kernel against the same
references in tests/test_convg_gpu.py)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

KINDS = {'s1': 0, 's2': 1, 't2': 2, 'p1s2': 3}


def describe(kind, role, n, h, w, cin0, cin1, cout, short):
    from tta_depth_completion_b200 import _lib
    L = _lib.lib()
    buf = (ctypes.c_int * 2048)()
    k = L.ptta_convg_plan_describe(KINDS[kind], role, n, h, w, cin0, cin1, cout, int(short), buf, 2048)
    assert k > 0, _lib.last_error()
    v = list(buf[:k])
    names = ('n_items', 'n_classes', 'th', 'tw', 'tiles_y', 'tiles_x', 'n_tiles', 'BN', 'halo', 'b_resident', 'n_a', 'n_b', 'in_parity',
             'out_parity', 'n_out', 'halo_rev')
    plan = dict(zip(names, v[:16]))
    plan['classes'] = [tuple(v[16 + 4 * c:20 + 4 * c]) for c in range(4)]
    plan['items'] = [tuple(v[32 + 8 * i:40 + 8 * i]) for i in range(plan['n_items'])]
    return plan


def replay(plan, kind, role, x0, x1, w, w_short, cin_w, cout_w):
    """out[n, gy*Q + qy, gx*Q + qx, :] = sum over the class's items of A_item[n, gy, gx, :] @ Wchunk_item^T"""
    P, Q, n_out = plan['in_parity'], plan['out_parity'], plan['n_out']
    srcs = [x0, x1]
    N, Hs, Ws = x0.shape[0], x0.shape[1], x0.shape[2]
    gh, gw = Hs // P, Ws // P
    out = torch.zeros((N, gh * Q, gw * Q, n_out), dtype=torch.float64)
    T = 1 if kind == 'p1s2' else 9
    conv_layout = kind != 't2'
    if role == 0:
        sn, sk, n_real, k_real = (cin_w * T, T, cout_w, cin_w) if conv_layout else (T, cout_w * T, cout_w, cin_w)
    else:
        sn, sk, n_real, k_real = (T, cin_w * T, cin_w, cout_w) if conv_layout else (cout_w * T, T, cin_w, cout_w)
    wf = w.reshape(-1).double()
    wsf = None if w_short is None else w_short.reshape(-1).double()
    for c in range(plan['n_classes']):
        start, count, out_c, out_py = plan['classes'][c]
        qy, qx = out_py, out_c // n_out
        acc = torch.zeros((N, gh, gw, n_out), dtype=torch.float64)
        for (c_inner, dx, dy, py, src, wsel, tap, k0) in plan['items'][start:start + count]:
            s = srcs[src]
            C = s.shape[3]
            px, c0 = divmod(c_inner, C)
            ys = (torch.arange(gh) + dy) * P + py
            xs = (torch.arange(gw) + dx) * P + px
            vy, vx = (ys >= 0) & (ys < Hs), (xs >= 0) & (xs < Ws)
            A = torch.zeros((N, gh, gw, 64), dtype=torch.float64)
            sub = s[:, ys[vy]][:, :, xs[vx]][..., c0:c0 + 64].double()
            iy, ix = torch.nonzero(vy).flatten(), torch.nonzero(vx).flatten()
            A[:, iy[:, None], ix[None, :], :] = sub
            Wc = torch.zeros((n_out, 64), dtype=torch.float64)
            nn_, kk = torch.arange(n_out), k0 + torch.arange(64)
            if wsel == 0:
                ok = (nn_[:, None] < n_real) & (kk[None, :] < k_real)
                idx = nn_[:, None] * sn + kk[None, :] * sk + tap
                Wc[ok] = wf[idx[ok]]
            else:                                # folded 1x1/s2 shortcut [cout][cin]: n = cin, k = cout
                ok = (nn_[:, None] < cin_w) & (kk[None, :] < cout_w)
                idx = nn_[:, None] + kk[None, :] * cin_w
                Wc[ok] = wsf[idx[ok]]
            acc += A @ Wc.t()
        out[:, qy::Q, qx::Q, :] = acc
    return out


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


CASES = [
    # kind, role, cin (stored; tuple for concat), cout, n, h, w, shortcut
    ('s1', 0, (64, 0), 64, 1, 5, 7, False),
    ('s1', 0, (64, 64), 128, 2, 4, 6, False),
    ('s1', 1, (128, 0), 64, 1, 5, 6, False),
    ('s2', 0, (64, 0), 128, 1, 6, 8, False),
    ('p1s2', 0, (64, 0), 128, 1, 6, 8, False),
    ('t2', 0, (128, 0), 64, 2, 3, 4, False),
    ('t2', 0, (64, 128), 64, 1, 3, 5, False),
    ('s2', 1, (64, 0), 128, 1, 6, 8, False),
    ('s2', 1, (64, 0), 128, 2, 4, 6, True),
    ('t2', 1, (192, 0), 64, 1, 3, 4, False),
]


@pytest.mark.parametrize('kind,role,cin,cout,n,h,w,short', CASES)
def test_plan_replay_equals_torch(kind, role, cin, cout, n, h, w, short):
    g = torch.Generator().manual_seed(sum(map(ord, kind)) + 17 * role + cout + h * w)
    c = cin[0] + cin[1]
    k = 1 if kind == 'p1s2' else 3
    wshape = (c, cout, 3, 3) if kind == 't2' else (cout, c, k, k)
    wt = torch.randn(wshape, generator=g, dtype=torch.float64)
    ws = torch.randn((cout, c, 1, 1), generator=g, dtype=torch.float64) if short else None
    plan = describe(kind, role, n, h, w, cin[0], cin[1], cout, short)

    def layer(x, weight, kd):
        if kd == 's1':
            return F.conv2d(x, weight, None, 1, 1)
        if kd == 's2':
            return F.conv2d(x, weight, None, 2, 1)
        if kd == 'p1s2':
            return F.conv2d(x, weight, None, 2, 0)
        return F.conv_transpose2d(x, weight, None, 2, 1, 1)

    if role == 0:
        x = torch.randn((n, c, h, w), generator=g, dtype=torch.float64)
        ref = layer(x, wt, kind)
        xs = nhwc(x)
        x0, x1 = xs[..., :cin[0]].contiguous(), (xs[..., cin[0]:].contiguous() if cin[1] else None)
        got = replay(plan, kind, role, x0, x1, wt, None, c, cout)
    else:
        x = torch.zeros((n, c, h, w), dtype=torch.float64, requires_grad=True)
        y = layer(x, wt, kind)
        gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        loss = (y * gy).sum()
        gys = None
        if short:
            ysh = layer(x, ws, 'p1s2')
            gys = torch.randn(ysh.shape, generator=g, dtype=torch.float64)
            loss = loss + (ysh * gys).sum()
        ref, = torch.autograd.grad(loss, x)
        got = replay(plan, kind, role, nhwc(gy), None if gys is None else nhwc(gys), wt, ws, c, cout)
    assert got.shape == nhwc(ref).shape
    assert float((got - nhwc(ref)).abs().max()) < 1e-9 * max(1.0, float(ref.abs().max()))


def test_plan_geometry_choices():
    """ring / tile policy: stride-1 layers use halo tiles (16x8 positions, one A tile per 64-channel chunk); weights stay resident
    when they fit; N tile = largest multiple of 64 <= 256 dividing the output channels"""
    p = describe('s1', 0, 1, 352, 1216, 64, 0, 64, False)
    assert (p['halo'], p['b_resident'], p['th'], p['tw'], p['BN'], p['n_items']) == (1, 1, 16, 8, 64, 9)
    assert p['tiles_y'] * p['tiles_x'] == 22 * 152 and p['n_a'] >= 4
    p = describe('s1', 0, 1, 88, 304, 256, 0, 256, False)
    assert (p['halo'], p['b_resident'], p['BN'], p['n_items'], p['n_a']) == (1, 0, 256, 36, 2) and p['n_b'] >= 4
    p = describe('s1', 0, 1, 352, 1216, 64, 64, 192, False)
    assert (p['BN'], p['n_tiles'], p['n_items']) == (192, 1, 18)
    p = describe('t2', 0, 1, 44, 152, 256, 512, 128, False)
    assert (p['halo'], p['n_classes'], p['out_parity'], p['n_items']) == (0, 4, 2, 108)
    assert [c[1] for c in p['classes']] == [12, 24, 24, 48]          # 1, 2, 2, 4 taps x 12 chunks
    p = describe('s2', 1, 1, 352, 1216, 64, 0, 128, True)
    assert p['classes'][0][1] == 2 + 2 and p['n_out'] == 64          # centre tap (K = 128: 2 chunks) + the folded shortcut's 2 chunks
    p = describe('s1', 0, 1, 352, 1216, 192, 64, 16, False)           # thin heads: 16 output channels, resident weights
    assert (p['BN'], p['b_resident'], p['n_items']) == (16, 1, 36)
