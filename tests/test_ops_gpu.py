"""Per-operator parity of the CUDA kernels (through the C ABI) against plain PyTorch fp32 on the same inputs.

Tolerances: integer-valued outputs (validity mask, filtered sparse depth) bit-exact; fp32 kernels 1e-5;
bf16-storage kernels are compared after rounding the fp32 reference to bf16, within a few bf16 ulps
(the kernels accumulate in fp32, so the only differences are summation order and the final rounding)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import msgchn_oracle as O


@pytest.fixture(scope='module')
def ops():
    from tta_depth_completion_b200 import ops as _ops
    return _ops


DEV = 'cuda'


def nhwc(x):      # NCHW float -> NHWC bf16 cuda
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)


def nchw(x):      # NHWC bf16 cuda -> NCHW float cpu
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


def bf(x):
    return x.to(torch.bfloat16).float()


def assert_close_bf16(got, want, what, ulps=4.0):
    """|got - want| <= ulps * 2^-8 * max(|want|, scale) elementwise, scale = rms of want"""
    scale = float(want.pow(2).mean().sqrt())
    tol = ulps * 2.0 ** -8 * torch.maximum(want.abs(), torch.full_like(want, scale))
    err = (got - want).abs()
    bad = err > tol
    assert not bool(bad.any()), '%s: %d / %d elements off, max err %.4g (rms %.4g)' % (what, int(bad.sum()), bad.numel(),
                                                                                       float(err.max()), scale)


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(1, 64, 128), (2, 48, 80), (1, 37, 53), (1, 352, 1216)])
def test_outlier_removal_bit_exact(ops, shape):
    n, h, w = shape
    _, sparse, _ = O.synthetic_frame(3, 1, n, h, w, 'kitti')
    want_d, want_v = O.remove_outliers(sparse, O.validity_map(sparse))
    d, v = ops.outlier_removal(sparse.to(DEV))
    assert torch.equal(v.cpu(), want_v) and torch.equal(d.cpu(), want_d)
    assert int(want_v.sum()) < int((sparse > 0).sum()) or h < 40      # the filter really removed something


def test_outlier_removal_edge_cases(ops):
    z = torch.zeros(1, 1, 32, 32)
    d, v = ops.outlier_removal(z.to(DEV))                     # empty frame
    assert float(d.abs().sum()) == 0 and float(v.sum()) == 0
    full = torch.full((1, 1, 32, 32), 5.0)
    full[0, 0, 10, 10] = 20.0                                  # one gross outlier in a dense map
    want_d, want_v = O.remove_outliers(full, O.validity_map(full))
    d, v = ops.outlier_removal(full.to(DEV))
    assert torch.equal(v.cpu(), want_v) and torch.equal(d.cpu(), want_d) and float(v[0, 0, 10, 10]) == 0.0


@pytest.mark.parametrize('shape', [(1, 64, 128), (2, 48, 80), (1, 352, 1216)])
def test_pyramid(ops, shape):
    n, h, w = shape
    _, sparse, _ = O.synthetic_frame(2, 0, n, h, w, 'kitti')
    sparse = sparse * 1.5                                       # some values above the 80 m cap
    dc, d2, d4 = ops.pyramid(sparse.to(DEV), 80.0)
    want = torch.clamp(sparse, 0, 80.0)
    w2, w4 = O.pyramid(want)
    assert torch.equal(dc.cpu(), want)
    assert torch.allclose(d2.cpu(), w2, rtol=1e-5, atol=1e-6) and torch.allclose(d4.cpu(), w4, rtol=1e-5, atol=1e-6)
    assert torch.equal((d2.cpu() > 0), (w2 > 0)) and torch.equal((d4.cpu() > 0), (w4 > 0))


# ------------------------------------------------------------------------------------------------------------
def _conv_case(n, h, w, cin, cout, seed):
    g = torch.Generator().manual_seed(seed)
    x = bf(torch.randn((n, cin, h, w), generator=g))
    wt = bf(torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (9 * cin)) ** 0.5)
    b = torch.randn((cout,), generator=g) * 0.1
    return x, wt, b


@pytest.mark.parametrize('n,h,w,cin,cout', [(1, 32, 48, 32, 32), (2, 19, 37, 32, 32), (1, 16, 32, 32, 128), (1, 16, 32, 128, 32),
                                            (1, 3, 5, 32, 32), (1, 88, 304, 32, 32)])
def test_conv_s1_forward_relu_prologue(ops, n, h, w, cin, cout):
    x, wt, b = _conv_case(n, h, w, cin, cout, 1)
    want = bf(F.conv2d(F.relu(x), wt, b, padding=1))
    got = ops.conv3x3(nhwc(x), ops.pack_conv_weight(wt.to(DEV), 'conv_fwd'), b.to(DEV), ops.MODE_S1, ops.PRO_RELU)
    assert_close_bf16(nchw(got), want, 'conv s1')


@pytest.mark.parametrize('n,h,w', [(1, 32, 48), (2, 18, 38), (1, 6, 10), (1, 176, 608)])
def test_conv_s2_forward(ops, n, h, w):
    x, wt, b = _conv_case(n, h, w, 32, 32, 2)
    want = bf(F.conv2d(F.relu(x), wt, b, stride=2, padding=1))
    got = ops.conv3x3(nhwc(x), ops.pack_conv_weight(wt.to(DEV), 'conv_fwd'), b.to(DEV), ops.MODE_S2, ops.PRO_RELU)
    assert_close_bf16(nchw(got), want, 'conv s2')


@pytest.mark.parametrize('n,h,w', [(1, 16, 24), (2, 9, 19), (1, 3, 5), (1, 88, 304)])
def test_conv_transposed_forward(ops, n, h, w):
    g = torch.Generator().manual_seed(3)
    x = bf(torch.randn((n, 32, h, w), generator=g))
    wt = bf(torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5)          # [Cin, Cout, 3, 3]
    b = torch.randn((32,), generator=g) * 0.1
    want = bf(F.conv_transpose2d(F.relu(x), wt, b, stride=2, padding=1, output_padding=1))
    got = ops.conv3x3(nhwc(x), ops.pack_conv_weight(wt.to(DEV), 'convT_fwd'), b.to(DEV), ops.MODE_T2, ops.PRO_RELU)
    assert_close_bf16(nchw(got), want, 'convT')


def test_conv_data_gradients(ops):
    """dgrad of conv s1 / conv s2 / convT against autograd, with ReLU mask and accumulate."""
    g = torch.Generator().manual_seed(4)
    n, h, w = 2, 16, 24
    wt = bf(torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5)
    pre = bf(torch.randn((n, 32, h, w), generator=g))                                # saved pre-activation (mask source)
    extra = bf(torch.randn((n, 32, h, w), generator=g))
    # stride 1: y = conv(relu(pre))
    x = pre.clone().requires_grad_(True)
    y = F.conv2d(F.relu(x), wt, None, padding=1)
    gy = bf(torch.randn(y.shape, generator=g))
    y.backward(gy)
    want = bf(x.grad + extra)
    got = ops.conv3x3(nhwc(gy), ops.pack_conv_weight(wt.to(DEV), 'conv_dgrad_s1'), None, ops.MODE_S1, ops.PRO_NONE,
                      mask=nhwc(pre), mask_mode=ops.MASK_RELU, add=nhwc(extra))
    assert_close_bf16(nchw(got), want, 'dgrad s1')
    # stride 2: y = conv_s2(relu(pre)) -> data gradient is a transposed conv
    x = pre.clone().requires_grad_(True)
    y = F.conv2d(F.relu(x), wt, None, stride=2, padding=1)
    gy = bf(torch.randn(y.shape, generator=g))
    y.backward(gy)
    got = ops.conv3x3(nhwc(gy), ops.pack_conv_weight(wt.to(DEV), 'conv_dgrad_s2'), None, ops.MODE_T2, ops.PRO_NONE,
                      mask=nhwc(pre), mask_mode=ops.MASK_RELU)
    assert_close_bf16(nchw(got), bf(x.grad), 'dgrad s2')
    # transposed: y = convT(relu(pre)) -> data gradient is a stride-2 conv
    x = pre.clone().requires_grad_(True)
    y = F.conv_transpose2d(F.relu(x), wt, None, stride=2, padding=1, output_padding=1)
    gy = bf(torch.randn(y.shape, generator=g))
    y.backward(gy)
    got = ops.conv3x3(nhwc(gy), ops.pack_conv_weight(wt.to(DEV), 'convT_dgrad'), None, ops.MODE_S2, ops.PRO_NONE,
                      mask=nhwc(pre), mask_mode=ops.MASK_RELU)
    assert_close_bf16(nchw(got), bf(x.grad), 'dgrad convT')


def test_conv_bn_leaky_prologue_and_mask(ops):
    g = torch.Generator().manual_seed(5)
    n, h, w = 1, 16, 32
    hraw = bf(torch.randn((n, 128, h, w), generator=g))
    scale = torch.rand((128,), generator=g) + 0.5
    shift = torch.randn((128,), generator=g) * 0.3
    wt = bf(torch.randn((32, 128, 3, 3), generator=g) * (2.0 / 1152) ** 0.5)
    b = torch.randn((32,), generator=g) * 0.1
    x = hraw.clone().requires_grad_(True)
    act = F.leaky_relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.2)
    y = F.conv2d(bf(act.detach()) + (act - act.detach()), wt, b, padding=1)       # forward sees the bf16-rounded activation
    got = ops.conv3x3(nhwc(hraw), ops.pack_conv_weight(wt.to(DEV), 'conv_fwd'), b.to(DEV), ops.MODE_S1, ops.PRO_BN_LEAKY,
                      pro_scale=scale.to(DEV), pro_shift=shift.to(DEV))
    assert_close_bf16(nchw(got), bf(y.detach()), 'conv bn+leaky prologue')
    # data gradient down to the BN output: g * leaky'(bn(h))
    gy = bf(torch.randn(y.shape, generator=g))
    act2 = F.leaky_relu(hraw * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), 0.2).requires_grad_(True)
    F.conv2d(act2, wt, b, padding=1).backward(gy)
    pre = hraw * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    want = bf(act2.grad * torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, 0.2)))
    got = ops.conv3x3(nhwc(gy), ops.pack_conv_weight(wt.to(DEV), 'conv_dgrad_s1'), None, ops.MODE_S1, ops.PRO_NONE,
                      mask=nhwc(hraw), mask_mode=ops.MASK_BN_LEAKY, mask_scale=scale.to(DEV), mask_shift=shift.to(DEV))
    assert_close_bf16(nchw(got), want, 'dgrad with leaky-bn mask')


@pytest.mark.parametrize('n,h,w,cin,cout', [(1, 16, 32, 32, 128), (2, 19, 21, 128, 32), (1, 22, 76, 32, 32)])
def test_conv_wgrad(ops, n, h, w, cin, cout):
    g = torch.Generator().manual_seed(6)
    x = bf(torch.randn((n, cin, h, w), generator=g))
    gy = bf(torch.randn((n, cout, h, w), generator=g))
    wt = torch.zeros((cout, cin, 3, 3), requires_grad=True)
    F.conv2d(x, wt, None, padding=1).backward(gy)
    got = ops.conv3x3_wgrad(nhwc(x), nhwc(gy)).cpu()
    rel = float((got - wt.grad).norm() / wt.grad.norm())
    assert rel < 2e-5, rel


def test_stem_and_head_conv(ops):
    g = torch.Generator().manual_seed(7)
    n, h, w = 2, 24, 40
    for cin in (1, 2, 3):
        x = torch.randn((n, cin, h, w), generator=g)
        wt = torch.randn((32, cin, 3, 3), generator=g) * 0.3
        b = torch.randn((32,), generator=g) * 0.1
        want = bf(F.conv2d(x, wt, b, padding=1))
        planes = [x[:, c].contiguous().to(DEV) for c in range(cin)]
        got = ops.stem_conv(planes, wt.to(DEV), b.to(DEV))
        assert_close_bf16(nchw(got), want, 'stem cin=%d' % cin, ulps=2.0)
        if w % 2 == 0:        # the engine's form: weights by value (constant bank); same FMA order -> bit-identical
            assert torch.equal(ops.stem_conv_const(planes, wt, b), got), 'stem_conv_const cin=%d' % cin
    x = bf(torch.randn((n, 32, h, w), generator=g))
    wt = torch.randn((1, 32, 3, 3), generator=g) * 0.1
    add = torch.randn((n, h, w), generator=g)
    want = F.conv2d(F.relu(x), wt, torch.tensor([0.25]), padding=1)[:, 0] + add
    w9 = wt[0].reshape(32, 9).t().contiguous()               # [tap][c]
    got = ops.head_conv(nhwc(x), w9.to(DEV), 0.25, add.to(DEV), relu_in=True)
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4)
    got = ops.head_conv_const(nhwc(x), w9, 0.25, add.to(DEV), relu_in=True)
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4), 'head_conv_const'
    acc = add.to(DEV).clone()
    ops.head_conv_const(nhwc(x), w9, 0.25, None, relu_in=True, out=acc, accumulate=True)
    assert torch.allclose(acc.cpu(), want, rtol=1e-4, atol=1e-4), 'head_conv_const accumulate'


@pytest.mark.parametrize('n,h,w', [(1, 8, 12), (2, 11, 19), (1, 37, 258), (1, 88, 304), (1, 352, 1216)])
@pytest.mark.parametrize('cin', [1, 2, 3])
def test_stem_conv_tc(ops, cin, n, h, w):
    """{1,2,3} -> 32 stem on tcgen05 (stem_tc_kernel: thread-built bf16 head + remainder operand tiles, six MMAs per 128 pixels, bulk store)
    vs the fp32 torch conv and vs the CUDA-core kernel it replaces; odd widths, partial last tile, multi-tile CTAs, scale / shift folding,
    and the masked no-bias form the prediction layers' data gradient uses."""
    g = torch.Generator().manual_seed(31 + cin + h)
    x = torch.randn((n, cin, h, w), generator=g) * 20.0                       # depth-like magnitudes: the head/remainder split has to hold
    wt = torch.randn((32, cin, 3, 3), generator=g) * 0.3
    b = torch.randn((32,), generator=g) * 0.1
    scale, shift = [0.5, 2.0, 1.0][:cin], [0.25, -1.0, 0.0][:cin]
    xn = x * torch.tensor(scale).view(1, cin, 1, 1) + torch.tensor(shift).view(1, cin, 1, 1)
    want = F.relu(F.conv2d(xn, wt, b, padding=1))
    planes = [x[:, c].contiguous().to(DEV) for c in range(cin)]
    got = ops.stem_conv_tc(planes, wt.to(DEV), b.to(DEV), scale=scale, shift=shift, relu_out=True)
    assert_close_bf16(nchw(got), bf(want), 'stem_tc cin=%d' % cin, ulps=2.0)
    old = ops.stem_conv(planes, wt.to(DEV), b.to(DEV), scale=scale, shift=shift)
    assert_close_bf16(nchw(got), torch.relu(nchw(old).float()), 'stem_tc vs stem cin=%d' % cin, ulps=2.0)
    if cin == 1:                                                             # data-gradient form: mask, no bias, no ReLU
        mask = bf(torch.randn((n, 32, h, w), generator=g))
        want = F.conv2d(x, wt, None, padding=1) * (mask > 0).float()
        got = ops.stem_conv_tc(planes, wt.to(DEV), None, mask=nhwc(mask), relu_out=False)
        assert_close_bf16(nchw(got), bf(want), 'stem_tc masked', ulps=2.0)


@pytest.mark.parametrize('n,h,w', [(1, 8, 12), (2, 24, 40), (1, 37, 258), (2, 9, 600), (1, 88, 304), (1, 352, 1216)])
def test_head_conv_tc(ops, n, h, w):
    """32 -> 1 conv on tcgen05 (conv3x3_tc_head_kernel: the conv_tc producer / MMA schedule with N = 3 x 16 and weights carried as a
    bf16 head + bf16 remainder) vs the fp32 torch conv of the same bf16 input and vs the CUDA-core kernel it replaces; multi-strip,
    ragged-width and multi-CTA row ranges included."""
    g = torch.Generator().manual_seed(17 + h)
    x = bf(torch.relu(torch.randn((n, 32, h, w), generator=g)))
    wt = torch.randn((1, 32, 3, 3), generator=g) * 0.1
    add = torch.randn((n, h, w), generator=g)
    want = F.conv2d(x, wt, torch.tensor([0.25]), padding=1)[:, 0] + add
    w9 = wt[0].reshape(32, 9).t().contiguous()               # [tap][c]
    got = ops.head_conv_tc(nhwc(x), w9.to(DEV), 0.25, add.to(DEV))
    old = ops.head_conv(nhwc(x), w9.to(DEV), 0.25, add.to(DEV), relu_in=False)
    err = float((got.cpu() - want).abs().max())
    assert err < 2e-4, err                                   # 288 products of O(0.1): weights exact to 2^-17, fp32 accumulation
    assert float((got - old).abs().max()) < 2e-4
    got = ops.head_conv_tc(nhwc(x), w9.to(DEV), 0.0, None)   # no addend, no bias
    assert float((got.cpu() - (want - add - 0.25)).abs().max()) < 2e-4


@pytest.mark.parametrize('n,h,w', [(1, 8, 12), (2, 11, 19), (1, 88, 304)])
def test_up2_and_adjoints(ops, n, h, w):
    g = torch.Generator().manual_seed(8)
    a = torch.randn((n, h, w), generator=g)
    b = torch.randn((n, h, w), generator=g)
    c = torch.randn((n, 2 * h, 2 * w), generator=g)
    want = F.interpolate((a + b).unsqueeze(1), scale_factor=2, mode='bilinear', align_corners=True)[:, 0] + c
    got = ops.up2_1ch(a.to(DEV), b.to(DEV), c.to(DEV))
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5)
    # adjoint against autograd
    al = a.clone().requires_grad_(True)
    F.interpolate(al.unsqueeze(1), scale_factor=2, mode='bilinear', align_corners=True).backward(c.unsqueeze(1))
    got = ops.up2_1ch_adjoint(c.to(DEV))
    assert torch.allclose(got.cpu(), al.grad, rtol=1e-4, atol=1e-5)
    # 32-channel bf16 variants
    x = bf(torch.randn((n, 32, 2 * h, 2 * w), generator=g))
    half = bf(torch.randn((n, 32, h, w), generator=g))
    want = bf(x + F.interpolate(half, scale_factor=2, mode='bilinear', align_corners=True))
    got = ops.add_up2_c32(nhwc(x), nhwc(half))
    assert_close_bf16(nchw(got), want, 'add_up2_c32', ulps=2.0)
    hl = half.clone().requires_grad_(True)
    F.interpolate(hl, scale_factor=2, mode='bilinear', align_corners=True).backward(x)
    got = ops.up2_c32_adjoint(nhwc(x))
    assert_close_bf16(nchw(got), bf(hl.grad), 'up2_c32_adjoint', ulps=2.0)


@pytest.mark.parametrize('m,n,k', [(300, 512, 32), (1000, 512, 512), (513, 32, 512), (26752, 512, 512)])
def test_gemm(ops, m, n, k):
    g = torch.Generator().manual_seed(9)
    a = bf(torch.randn((m, k), generator=g))
    b = bf(torch.randn((n, k), generator=g) / k ** 0.5)
    bias = torch.randn((n,), generator=g)
    want = bf(a.to(DEV) @ b.to(DEV).t() + bias.to(DEV)).cpu()      # fp32 matmul of bf16-exact inputs (TF32 off by default)
    got = ops.gemm_bf16(a.to(DEV, torch.bfloat16), b.to(DEV, torch.bfloat16), bias.to(DEV)).float().cpu()
    assert_close_bf16(got, want, 'gemm %dx%dx%d' % (m, n, k))


@pytest.mark.parametrize('m,n,k', [(300, 256, 64), (1000, 512, 512), (26752, 512, 512), (129, 768, 128), (26752, 512, 32), (97, 512, 32), (40000, 256, 64)])
def test_gemm_tcgen05(ops, m, n, k):
    g = torch.Generator().manual_seed(11)
    a = bf(torch.randn((m, k), generator=g))
    b = bf(torch.randn((n, k), generator=g) / k ** 0.5)
    bias = torch.randn((n,), generator=g)
    want = bf(a.to(DEV) @ b.to(DEV).t() + bias.to(DEV)).cpu()
    got = ops.gemm_bf16_tc(a.to(DEV, torch.bfloat16), b.to(DEV, torch.bfloat16), bias.to(DEV)).float().cpu()
    assert_close_bf16(got, want, 'gemm_tc %dx%dx%d' % (m, n, k))
    got2 = ops.gemm_bf16(a.to(DEV, torch.bfloat16), b.to(DEV, torch.bfloat16), bias.to(DEV)).float().cpu()
    assert float((got - got2).abs().max()) <= 2.0 ** -6 * float(want.abs().max())     # both accumulate in fp32


@pytest.mark.parametrize('n,h,w', [(1, 16, 128), (2, 24, 300), (1, 88, 304), (1, 352, 1216)])
def test_conv_tcgen05_fused_upsampled_addend(ops, n, h, w):
    """x = conv(a) + up2(pre) in ONE pass (epilogue of conv3x3_tc_kernel gathers the four half-resolution neighbours) vs torch
    (F.conv2d + F.interpolate align_corners=True on the same bf16 inputs) and vs the two-pass form (conv, then add_up2_c32)"""
    g = torch.Generator().manual_seed(23 + h)
    x = bf(torch.relu(torch.randn((n, 32, h, w), generator=g)))
    half = bf(torch.randn((n, 32, h // 2, w // 2), generator=g))
    wt = torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5
    b = torch.randn((32,), generator=g) * 0.1
    want = F.conv2d(x, bf(wt), b, padding=1) + F.interpolate(half, scale_factor=2, mode='bilinear', align_corners=True)
    wp = ops.pack_conv_weight(wt.to(DEV), 'conv_fwd')
    out, out_relu = ops.conv3x3_tc_up2(nhwc(x), wp, nhwc(half), b.to(DEV))
    assert_close_bf16(nchw(out), bf(want), 'conv_tc + up2', ulps=2.0)
    assert torch.equal(out_relu, torch.relu(out.float()).to(torch.bfloat16))
    two = ops.add_up2_c32(ops.conv3x3_tc(nhwc(x), wp, b.to(DEV)), nhwc(half))       # rounds twice (<= 2 x 2^-8 each way) vs once
    assert_close_bf16(nchw(out), nchw(two), 'fused vs two-pass', ulps=3.0)


@pytest.mark.parametrize('n,h,w', [(1, 16, 128), (1, 24, 300), (2, 19, 38), (1, 88, 304), (3, 64, 514), (1, 352, 1216)])
def test_conv_tcgen05_matches_mma(ops, n, h, w):
    """the tcgen05 conv is bit-identical to the mma.sync kernel (same bf16 products, fp32 accumulation of 288 terms)"""
    g = torch.Generator().manual_seed(12)
    wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(DEV)
    wp = ops.pack_conv_weight(wt, 'conv_fwd')
    bias = (torch.randn(32, generator=g) * 0.1).to(DEV)
    x = torch.randn((n, h, w, 32), generator=g).to(DEV).to(torch.bfloat16)
    m = torch.randn((n, h, w, 32), generator=g).to(DEV).to(torch.bfloat16)
    a = torch.randn((n, h, w, 32), generator=g).to(DEV).to(torch.bfloat16)
    # producers store ReLU(x) (relu_out), so the tcgen05 kernel has no ReLU-on-load: feed it the ReLU'd map
    want = ops.conv3x3(x, wp, bias, ops.MODE_S1, ops.PRO_RELU)
    got = ops.conv3x3_tc(torch.relu(x), wp, bias)
    assert torch.equal(got, want), 'conv_tc forward is not bit-identical to mma.sync'
    got = ops.conv3x3_tc(torch.relu(x), wp, bias, relu_out=True)
    assert torch.equal(got, torch.relu(want)), 'conv_tc relu_out'
    want = ops.conv3x3(x, wp, None, ops.MODE_S1, ops.PRO_NONE, mask=m, mask_mode=ops.MASK_RELU, add=a)
    got = ops.conv3x3_tc(x, wp, None, mask=m, add=a)
    assert torch.equal(got, want), 'conv_tc mask+add is not bit-identical to mma.sync'
    acc = a.clone()                                    # in-place accumulate (out aliases add), as the backward pass uses it
    from tta_depth_completion_b200 import _lib
    wi = ops.pack_conv_weight_tc(wp)
    _lib.check(_lib.lib().ptta_conv3x3_tc(_lib.ptr(x), _lib.ptr(acc), _lib.ptr(wi), None, n, h, w, 0, 0, _lib.ptr(m), _lib.ptr(acc),
                                          torch.cuda.current_stream().cuda_stream), 'conv3x3_tc in place')
    assert torch.equal(acc, want), 'conv_tc in-place accumulate'
    # second output: ReLU copy, and the decoder sum ReLU(bf16(out) + skip)
    base = ops.conv3x3_tc(torch.relu(x), wp, bias)
    o, o2 = ops.conv3x3_tc_ex(torch.relu(x), wp, bias)
    assert torch.equal(o, base) and torch.equal(o2, torch.relu(base)), 'conv_tc out2 = ReLU(out)'
    o, o2 = ops.conv3x3_tc_ex(torch.relu(x), wp, bias, add2=a)
    assert torch.equal(o, base) and torch.equal(o2, torch.relu((base.float() + a.float()).to(torch.bfloat16))), 'conv_tc out2 = ReLU(out + add2)'
    # independent of the mma.sync kernel: fp32 convolution of the same bf16 operands (fp32 accumulation of 288 products; the only
    # differences are summation order and the final rounding to bf16)
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(torch.relu(x).float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, 1, 1).permute(0, 2, 3, 1)
    err = (base.float() - ref).abs()
    assert int((err > ref.abs() * 2.0 ** -8 + 2e-3 * float(ref.pow(2).mean().sqrt())).sum()) == 0, 'conv_tc vs fp32 conv'
    with pytest.raises(RuntimeError):
        ops.conv3x3_tc(x, wp, bias, relu_in=True)      # not supported: must fail loudly
    with pytest.raises(RuntimeError):
        ops.conv3x3_tc(x[:, :, :w - 1].contiguous(), wp, bias)      # odd width


@pytest.mark.parametrize('n,h,w', [(1, 16, 256), (1, 32, 48), (2, 18, 38), (1, 88, 304), (3, 64, 516), (1, 352, 1216)])
def test_conv_tcgen05_stride2_matches_mma(ops, n, h, w):
    """the stride-2 tcgen05 conv (Conv2d(s2) forward / ConvTranspose2d(s2) data gradient) is bit-identical to the mma.sync kernel"""
    g = torch.Generator().manual_seed(13)
    wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(DEV)
    wp = ops.pack_conv_weight(wt, 'conv_fwd')
    bias = (torch.randn(32, generator=g) * 0.1).to(DEV)
    x = torch.randn((n, h, w, 32), generator=g).to(DEV).to(torch.bfloat16)
    m = torch.randn((n, h // 2, w // 2, 32), generator=g).to(DEV).to(torch.bfloat16)
    a = torch.randn((n, h // 2, w // 2, 32), generator=g).to(DEV).to(torch.bfloat16)
    want = ops.conv3x3(x, wp, bias, ops.MODE_S2, ops.PRO_RELU)
    got, got_relu = ops.conv3x3_tc_s2(torch.relu(x), wp, bias, want_relu_copy=True)
    assert torch.equal(got, want), 'conv_tc_s2 forward is not bit-identical to mma.sync'
    assert torch.equal(got_relu, torch.relu(want)), 'conv_tc_s2 ReLU copy'
    assert torch.equal(ops.conv3x3_tc_s2(torch.relu(x), wp, bias, relu_out=True), torch.relu(want))
    want = ops.conv3x3(x, wp, None, ops.MODE_S2, ops.PRO_NONE, mask=m, mask_mode=ops.MASK_RELU, add=a)
    assert torch.equal(ops.conv3x3_tc_s2(x, wp, None, mask=m, add=a), want), 'conv_tc_s2 mask+add'
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(torch.relu(x).float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, 2, 1).permute(0, 2, 3, 1)
    err = (got.float() - ref).abs()
    assert int((err > ref.abs() * 2.0 ** -8 + 2e-3 * float(ref.pow(2).mean().sqrt())).sum()) == 0, 'conv_tc_s2 vs fp32 conv'
    with pytest.raises(RuntimeError):
        ops.conv3x3_tc_s2(x[:, :h - 1].contiguous(), wp, bias)       # odd height: must fail loudly


@pytest.mark.parametrize('n,h,w', [(1, 8, 128), (1, 16, 24), (2, 9, 38), (1, 3, 6), (1, 44, 152), (3, 33, 258), (1, 88, 304), (1, 176, 608)])
def test_conv_tcgen05_transposed_matches_mma_and_torch(ops, n, h, w):
    """the transposed stride-2 tcgen05 conv (ConvTranspose2d(s2) forward / Conv2d(s2) data gradient): equal to the mma.sync kernel up to
    the summation order of the 4-tap output class (a fraction < 1e-3 of the elements may differ by one bf16 ulp; measured 2e-4), and
    within bf16 rounding of an fp32 conv_transpose2d of the same bf16 operands"""
    def same(a, b, what):
        d = (a.float() - b.float()).abs()
        assert float((d > 0).float().mean()) < 1e-3 and bool((d <= b.float().abs() * 2.0 ** -7 + 1e-6).all()), what
    g = torch.Generator().manual_seed(14)
    wt = (torch.randn((32, 32, 3, 3), generator=g) * (2.0 / 288) ** 0.5).to(DEV)          # ConvTranspose2d weight [Cin, Cout, 3, 3]
    wp = ops.pack_conv_weight(wt, 'convT_fwd')
    bias = (torch.randn(32, generator=g) * 0.1).to(DEV)
    x = torch.randn((n, h, w, 32), generator=g).to(DEV).to(torch.bfloat16)
    m = torch.randn((n, 2 * h, 2 * w, 32), generator=g).to(DEV).to(torch.bfloat16)
    a = torch.randn((n, 2 * h, 2 * w, 32), generator=g).to(DEV).to(torch.bfloat16)
    want = ops.conv3x3(x, wp, bias, ops.MODE_T2, ops.PRO_RELU)
    got = ops.conv3x3_tc_t2(torch.relu(x), wp, bias)
    same(got, want, 'conv_tc_t2 forward vs mma.sync')
    assert torch.equal(ops.conv3x3_tc_t2(torch.relu(x), wp, bias, relu_out=True), torch.relu(got)), 'conv_tc_t2 relu_out'
    want2 = ops.conv3x3(x, wp, None, ops.MODE_T2, ops.PRO_NONE, mask=m, mask_mode=ops.MASK_RELU, add=a)
    same(ops.conv3x3_tc_t2(x, wp, None, mask=m, add=a), want2, 'conv_tc_t2 mask+add vs mma.sync')
    acc = a.clone()                                    # in-place accumulate (out aliases add), as the backward pass uses it
    from tta_depth_completion_b200 import _lib
    wi = torch.empty((9 * 32 * 32,), dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().ptta_pack_conv_weight_tc_t2(_lib.ptr(wp), _lib.ptr(wi), torch.cuda.current_stream().cuda_stream), 'pack')
    _lib.check(_lib.lib().ptta_conv3x3_tc_t2(_lib.ptr(x), _lib.ptr(acc), _lib.ptr(wi), None, n, h, w, 0, _lib.ptr(m), _lib.ptr(acc),
                                             torch.cuda.current_stream().cuda_stream), 'conv3x3_tc_t2 in place')
    same(acc, want2, 'conv_tc_t2 in-place accumulate')
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv_transpose2d(torch.relu(x).float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, stride=2, padding=1,
                             output_padding=1).permute(0, 2, 3, 1)
    err = (got.float() - ref).abs()
    assert int((err > ref.abs() * 2.0 ** -8 + 2e-3 * float(ref.pow(2).mean().sqrt())).sum()) == 0, 'conv_tc_t2 vs fp32 conv_transpose2d'
    with pytest.raises(RuntimeError):
        ops.conv3x3_tc_t2(x[:, :, :w - 1].contiguous(), wp, bias)       # odd width: must fail loudly


def test_adam_matches_torch(ops):
    g = torch.Generator().manual_seed(10)
    p0 = torch.randn((5000,), generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p = p0.clone().to(DEV)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn((5000,), generator=g) * 10.0 ** float(torch.randint(-6, 1, (1,), generator=g))
        ref.grad = grad.clone()
        opt.step()
        ops.adam_flat(p, grad.to(DEV), m, v, 1e-3, step)
    assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-6, atol=1e-7)
    assert torch.allclose(m.cpu(), opt.state[ref]['exp_avg'], rtol=1e-6, atol=1e-12)
    assert torch.allclose(v.cpu(), opt.state[ref]['exp_avg_sq'], rtol=1e-6, atol=1e-20)


@pytest.mark.gpu
def test_input_stage_matches_numpy_loader():
    """ptta_input_stage == src/data_utils.py:134-200 (np.asarray(image, float32) transposed to CHW; depth = png / 256, <= 0 -> 0,
    validity = depth > 0) followed by the bottom / centre crop of src/datasets.py:83-170, bit-exact"""
    import numpy as np
    from tta_depth_completion_b200 import ops
    rng = np.random.RandomState(3)
    n, h0, w0, h, w = 2, 375, 1242, 352, 1216
    img = rng.randint(0, 256, size=(n, h0, w0, 3)).astype(np.uint8)
    png = (rng.randint(0, 65536, size=(n, h0, w0)) * (rng.rand(n, h0, w0) < 0.05)).astype(np.uint16)
    # reference arithmetic
    image_ref = np.transpose(np.asarray(img, np.float32), (0, 3, 1, 2))
    z = png.astype(np.float32) / 256.0
    z[z <= 0] = 0.0
    v = z.astype(np.float32).copy()
    v[z > 0] = 1.0
    x0, y0 = (w0 - w) // 2, h0 - h
    image_ref, z, v = image_ref[:, :, y0:y0 + h, x0:x0 + w], z[:, None, y0:y0 + h, x0:x0 + w], v[:, None, y0:y0 + h, x0:x0 + w]
    dev = torch.device('cuda:0')
    image, depth, validity = ops.input_stage(torch.from_numpy(img).to(dev), torch.from_numpy(png.view(np.int16)).to(dev), (h, w), ('bottom',))
    assert torch.equal(image.cpu(), torch.from_numpy(np.ascontiguousarray(image_ref)))
    assert torch.equal(depth.cpu(), torch.from_numpy(np.ascontiguousarray(z)))
    assert torch.equal(validity.cpu(), torch.from_numpy(np.ascontiguousarray(v)))
    with pytest.raises(RuntimeError):
        ops.input_stage(torch.from_numpy(img).to(dev), torch.from_numpy(png.view(np.int16)).to(dev), (400, 1216))
