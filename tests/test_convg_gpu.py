"""General-channel tcgen05 convolution family (csrc/conv_gen.cuh) against PyTorch fp32 convolutions of the same bf16-rounded
operands: every layer geometry of the NLSPN network (nlspnmodel_adapt.py:384-448) in the forward and data-gradient roles.
Tolerance: the kernel accumulates in fp32 and stores bf16, so the only differences are summation order and the final
rounding: |out - ref| <= 2^-8 |ref| + 2e-3 * rms(ref) element-wise, and <= 3e-3 norm-wise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _seed(*a):
    return sum((i + 1) * 7919 * (sum(map(ord, str(v))) % 1009) for i, v in enumerate(a)) & 0xffffff


def _dev():
    return torch.device('cuda:0')


def _rand_nhwc(g, n, h, w, c, dev):
    x = torch.randn((n, h, w, c), generator=g, dtype=torch.float32).to(dev).to(torch.bfloat16)
    return x


def _nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def _check(out, ref_nchw, what):
    ref = ref_nchw.permute(0, 2, 3, 1)
    got = out.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    rms = float(ref.pow(2).mean().sqrt())
    err = (got - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + 2e-3 * rms
    bad = int((err > bound).sum())
    nrel = float((got - ref).norm() / ref.norm())
    assert bad == 0 and nrel < 3e-3, '%s: %d elements out of bound, norm-wise %.2e, max err %.3e (rms %.3e)' % (what, bad, nrel, float(err.max()), rms)


def _layer(kind, wt, x, stride_pad=None):
    if kind == 's1':
        return F.conv2d(x, wt, None, 1, 1)
    if kind == 's2':
        return F.conv2d(x, wt, None, 2, 1)
    if kind == 'p1s2':
        return F.conv2d(x, wt, None, 2, 0)
    return F.conv_transpose2d(x, wt, None, 2, 1, 1)


FWD_CASES = [
    # kind, (cin0, cin1), cout, n, h, w
    ('s1', (64, 0), 64, 1, 24, 40),
    ('s1', (64, 0), 64, 2, 17, 23),
    ('s1', (128, 0), 128, 1, 16, 48),
    ('s1', (64, 64), 64, 1, 20, 36),
    ('s1', (512, 0), 512, 1, 6, 10),
    ('s1', (256, 0), 256, 2, 8, 20),
    ('s2', (64, 0), 128, 1, 32, 48),
    ('s2', (256, 0), 512, 2, 12, 20),
    ('p1s2', (64, 0), 128, 1, 32, 48),
    ('p1s2', (128, 0), 256, 2, 10, 36),
    ('t2', (512, 0), 256, 2, 3, 5),
    ('t2', (256, 512), 128, 1, 6, 10),
    ('t2', (64, 128), 64, 1, 24, 40),
    ('s1', (64, 0), 64, 1, 1, 130),
    ('s1', (64, 0), 64, 1, 40, 200),
    # multi-wave launches (tiles >> 148 CTAs): the persistent `tile += gridDim.x` loops, ring slot / phase wrap, the TMEM
    # accumulator ring and the second issuing thread -- the code paths every full-size NLSPN number comes from
    ('s1', (64, 0), 64, 1, 352, 1216),      # resnet34.layer1: 3 344 tiles, resident weights, two issuing threads
    ('s1', (128, 0), 128, 1, 176, 608),     # layer2: streamed weights
    ('s1', (256, 0), 256, 1, 88, 304),      # layer3
    ('s1', (512, 0), 512, 1, 44, 152),      # layer4
    ('s1', (64, 64), 192, 1, 352, 1216),    # id/gd/cf_dec1 as one 128->192 conv over a concat
    ('s2', (64, 0), 128, 1, 352, 1216),     # layer2.0 stride 2
    ('p1s2', (64, 0), 128, 1, 352, 1216),   # layer2.0 downsample
    ('t2', (64, 128), 64, 1, 176, 608),     # dec2: ConvTranspose 192->64 over a skip concat
    ('t2', (128, 256), 64, 1, 88, 304),     # dec3
    ('s1', (64, 0), 64, 2, 180, 612),       # batch 2, ragged tiles on both edges
]


@pytest.mark.parametrize('kind,cin,cout,n,h,w', FWD_CASES)
def test_forward(kind, cin, cout, n, h, w):
    from tta_depth_completion_b200.convg import ConvG, FWD
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = _dev()
    g = torch.Generator().manual_seed(_seed(kind, cin, cout, n, h, w))
    c = cin[0] + cin[1]
    k = 1 if kind == 'p1s2' else 3
    shape = (c, cout, 3, 3) if kind == 't2' else (cout, c, k, k)
    wt = (torch.randn(shape, generator=g) * (2.0 / (c * k * k)) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev) * 0.1
    x0 = _rand_nhwc(g, n, h, w, cin[0], dev)
    x1 = _rand_nhwc(g, n, h, w, cin[1], dev) if cin[1] else None
    op = ConvG(kind, FWD, wt, cin if cin[1] else cin[0], cout, bias=bias)
    out = op(x0, x1)
    xin = _nchw(x0) if x1 is None else torch.cat((_nchw(x0), _nchw(x1)), 1)
    ref = _layer(kind, wt.to(torch.bfloat16).float(), xin) + bias.view(1, -1, 1, 1)
    _check(out, ref, 'fwd %s %s->%d %dx%dx%d' % (kind, cin, cout, n, h, w))


DGRAD_CASES = [
    # kind, cin, cout, n, h, w (layer input size), with_short
    ('s1', 64, 64, 1, 24, 40, False),
    ('s1', 128, 64, 2, 17, 23, False),
    ('s1', 512, 512, 1, 6, 10, False),
    ('s2', 64, 128, 1, 32, 48, False),
    ('s2', 64, 128, 2, 16, 40, True),
    ('s2', 256, 512, 1, 12, 20, True),
    ('s2', 512, 512, 1, 6, 10, False),
    ('t2', 768, 128, 1, 6, 10, False),
    ('t2', 192, 64, 1, 24, 40, False),
    ('t2', 512, 256, 2, 3, 5, False),
    # multi-wave (layer sizes of the 352x1216 step)
    ('s1', 64, 64, 1, 352, 1216, False),
    ('s1', 128, 128, 1, 176, 608, False),
    ('s2', 64, 128, 1, 352, 1216, True),    # stride-2 data gradient with the folded 1x1/s2 shortcut, 4 parity classes
    ('s2', 128, 256, 1, 176, 608, True),
    ('t2', 192, 64, 1, 176, 608, False),
    ('s1', 64, 64, 2, 180, 612, False),
]


@pytest.mark.parametrize('kind,cin,cout,n,h,w,short', DGRAD_CASES)
def test_data_gradient(kind, cin, cout, n, h, w, short):
    from tta_depth_completion_b200.convg import ConvG, DGRAD
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = _dev()
    g = torch.Generator().manual_seed(_seed(kind, cin, cout, n, h, w, short))
    shape = (cin, cout, 3, 3) if kind == 't2' else (cout, cin, 3, 3)
    wt = (torch.randn(shape, generator=g) * (2.0 / (cin * 9)) ** 0.5).to(dev)
    ws = (torch.randn((cout, cin, 1, 1), generator=g) * (2.0 / cin) ** 0.5).to(dev) if short else None
    x = torch.zeros((n, cin, h, w), device=dev, requires_grad=True)
    y = _layer(kind, wt.to(torch.bfloat16).float(), x)
    gy = _rand_nhwc(g, n, y.shape[2], y.shape[3], cout, dev)
    loss = (y * _nchw(gy)).sum()
    gys = None
    if short:
        ys = _layer('p1s2', ws.to(torch.bfloat16).float(), x)
        gys = _rand_nhwc(g, n, ys.shape[2], ys.shape[3], cout, dev)
        loss = loss + (ys * _nchw(gys)).sum()
    ref, = torch.autograd.grad(loss, x)
    op = ConvG(kind, DGRAD, wt, cin, cout, weight_short=ws)
    out = op(gy, gys, hw=(h, w))
    _check(out, ref, 'dgrad %s %d->%d %dx%dx%d short=%s' % (kind, cin, cout, n, h, w, short))


def test_thin_head_conv_full_size():
    """id_dec0 | gd_dec0 | cf_dec0 as ONE 16-output-channel conv over (F | fe1) with fp32 planar outputs and a per-channel
    activation (nlspnmodel_adapt.py:430-448, 883-895), at 352x1216 (multi-wave) against fp32 convolutions."""
    import ctypes
    from tta_depth_completion_b200 import _lib
    from tta_depth_completion_b200._lib import check, ptr, c_void_p
    from tta_depth_completion_b200.convg import ConvG, FWD
    torch.backends.cudnn.allow_tf32 = False
    dev = _dev()
    for (n, h, w) in ((1, 352, 1216), (2, 40, 72)):
        g = torch.Generator().manual_seed(31 + h)
        wt = torch.zeros((16, 256, 3, 3))
        wt[:10] = torch.randn((10, 256, 3, 3), generator=g) * (2.0 / (256 * 9)) ** 0.5
        wt = wt.to(dev)
        bias = torch.zeros(16, device=dev)
        bias[:10] = (torch.randn(10, generator=g) * 0.1).to(dev)
        x0, x1 = _rand_nhwc(g, n, h, w, 192, dev), _rand_nhwc(g, n, h, w, 64, dev)
        op = ConvG('s1', FWD, wt, (192, 64), 16, bias=bias)
        out = torch.empty((10, n, h, w), dtype=torch.float32, device=dev)
        planes = (ctypes.c_void_p * 10)(*[out[c].data_ptr() for c in range(10)])
        strides = (ctypes.c_longlong * 10)(*([h * w] * 10))
        acts_l = [1, 0, 0, 0, 0, 0, 0, 0, 0, 2]              # LeakyReLU (init depth) | identity x 8 (guidance) | sigmoid (confidence)
        acts = (ctypes.c_int * 10)(*acts_l)
        check(_lib.lib().ptta_convg_run_thin(ptr(x0), ptr(x1), ptr(op.packed), ptr(op.bias), planes, strides, acts, 10, n, h, w, 192, 64,
                                             c_void_p(torch.cuda.current_stream().cuda_stream)), 'convg_run_thin')
        ref = F.conv2d(torch.cat((_nchw(x0), _nchw(x1)), 1), wt[:10].to(torch.bfloat16).float(), bias[:10], 1, 1)
        ref[:, 0] = F.leaky_relu(ref[:, 0], 0.2)
        ref[:, 9] = torch.sigmoid(ref[:, 9])
        got = out.permute(1, 0, 2, 3)
        e = float((got - ref).norm() / ref.norm())
        assert e < 2e-3, (n, h, w, e)
        assert float((got - ref).abs().max()) < 2e-2 * float(ref.abs().max()), (n, h, w)


def test_meta_conv_carries_depth_channels():
    """48->48 meta conv stored as 64->64: output channels 48..63 are bit-exact copies of input channels 48..63 (conv1_dep's
    features), so the kernel writes fe1 = cat(fe1_rgb, fe1_dep) directly (nlspnmodel_adapt.py:866-870)."""
    from tta_depth_completion_b200.convg import ConvG, FWD
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    wt = (torch.randn((48, 48, 3, 3), generator=g) * 0.05).to(dev)
    bias = (torch.randn(48, generator=g) * 0.1).to(dev)
    x = _rand_nhwc(g, 2, 19, 37, 64, dev)
    op = ConvG('s1', FWD, wt, 64, 64, bias=bias, ident_from=48)
    out = op(x)
    assert torch.equal(out[..., 48:], x[..., 48:])
    ref = F.conv2d(_nchw(x)[:, :48], wt.to(torch.bfloat16).float(), bias, 1, 1)
    _check(out[..., :48], ref, 'meta conv 48->48')


def test_weight_multicast_switch_is_bit_identical():
    """experiment switch 32 (ptta_convg_debug_set): the streamed weight tiles are loaded half by each CTA of a 2-CTA cluster and
    multicast to both; results must be bit-identical to the default launch (odd tile counts run a clipped ghost tile)"""
    from tta_depth_completion_b200 import _lib
    from tta_depth_completion_b200.convg import ConvG, FWD, DGRAD
    dev = _dev()
    g = torch.Generator().manual_seed(77)
    L = _lib.lib()
    try:
        for c, n, h, w, role in ((128, 1, 40, 72, FWD), (256, 2, 17, 23, FWD), (128, 1, 48, 40, DGRAD)):
            wt = (torch.randn((c, c, 3, 3), generator=g) * (2.0 / (c * 9)) ** 0.5).to(dev)
            x = _rand_nhwc(g, n, h, w, c, dev)
            op = ConvG('s1', role, wt, c, c)
            L.ptta_convg_debug_set(0)
            ref = op(x, hw=(h, w))
            L.ptta_convg_debug_set(32)
            got = op(x, hw=(h, w))
            torch.cuda.synchronize()
            assert torch.equal(ref, got), (c, n, h, w, role)
    finally:
        L.ptta_convg_debug_set(0)
