"""TEST INFRASTRUCTURE (checker only -- never imported by the product path).

CPU restatement of the NLSPN non-local spatial propagation path of seobbro/TTA-depth-completion (SURVEY.md section 8
rows a19-a21), in numpy / torch fp32 (optionally fp64):

  * `mdconv_forward` / `mdconv_backward`: the modulated deformable convolution (DCNv2) exactly as the reference's CUDA
    kernels compute it -- external_src/NLSPN/src/model/deformconv/src/cuda/modulated_deform_im2col_cuda.cuh:
      :24-54    mdmcn_im2col_bilinear        (corner validity rules, floor-based corners)
      :56-84    mdmcn_get_gradient_weight    (bilinear scatter weights of the input gradient)
      :86-126   mdmcn_get_coordinate_weight  (d/dh, d/dw of the bilinear sample)
      :128-194  modulated_deformable_im2col_gpu_kernel   (sample window (-1,H)x(-1,W), value * mask)
      :196-252  modulated_deformable_col2im_gpu_kernel   (grad_input scatter)
      :254-328  modulated_deformable_col2im_coord_gpu_kernel (grad_offset, grad_mask)
    and the host code modulated_deform_conv_cuda.cu:19-128 (columns x weight + bias), :131-290 (backward).
    Layouts are the reference's: input [N,C,H,W], offset [N, 2*K, Ho, Wo] ordered (dh, dw) per tap, mask [N, K, Ho, Wo],
    weight [Cout, Cin, kh, kw], groups = deformable_groups = 1.
  * `MDConvFn`: torch.autograd.Function with the argument list of the reference's `ModulatedDeformConvFunction`
    (external_src/NLSPN/src/model/modulated_deform_conv_func.py:15-56) backed by the two functions above.
  * `offset_affinity` and `propagate`: NLSPN._get_offset_affinity / NLSPN.forward
    (external_src/NLSPN/src/model/nlspnmodel_adapt.py:255-330, :340-373) for affinity 'TGASS', conf_prop, preserve_input.

Pinning (tests/test_nlspn_oracle.py, tests/golden/nlspn_prop_*.pt): (1) the reference's own NLSPN Python class is executed
in this container with its `ModulatedDeformConvFunction` import resolved to `MDConvFn` (the DCN CUDA extension has no CPU
path) and its outputs / gradients are stored as golden fixtures by oracle/gen_golden_nlspn.py; (2) `mdconv_*` agree with
torchvision.ops.deform_conv2d (same DCNv2 lineage, torchvision 0.26) and its autograd in fp64; (3) on the GPU box the
reference's own CUDA kernels, built unmodified-in-arithmetic by oracle/build_ref_dcn.py into oracle/_ref/, are compared
with both (tests/test_nlspn_gpu.py)."""
import numpy as np
import torch


def _geometry(H, W, kh, kw, stride, pad, dil):
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    return Ho, Wo


def _positions(offset, kh, kw, stride, pad, dil, Ho, Wo, dtype):
    """sampling positions h_im, w_im: [N, K, Ho, Wo] (modulated_deform_im2col_cuda.cuh:166-176)"""
    N = offset.shape[0]
    K = kh * kw
    hs = (np.arange(Ho, dtype=dtype) * stride - pad)[None, None, :, None]
    ws = (np.arange(Wo, dtype=dtype) * stride - pad)[None, None, None, :]
    ki = (np.arange(K) // kw).astype(dtype)[None, :, None, None] * dil
    kj = (np.arange(K) % kw).astype(dtype)[None, :, None, None] * dil
    off = offset.reshape(N, K, 2, Ho, Wo)
    h_im = hs + ki + off[:, :, 0]
    w_im = ws + kj + off[:, :, 1]
    return h_im.astype(dtype), w_im.astype(dtype)


def _corners(h_im, w_im, H, W):
    """floor corners, bilinear weights and corner validity (cuh:24-54)"""
    h_low = np.floor(h_im).astype(np.int64)
    w_low = np.floor(w_im).astype(np.int64)
    h_high, w_high = h_low + 1, w_low + 1
    lh = h_im - h_low
    lw = w_im - w_low
    hh, hw = 1 - lh, 1 - lw
    ok1 = (h_low >= 0) & (w_low >= 0)
    ok2 = (h_low >= 0) & (w_high <= W - 1)
    ok3 = (h_high <= H - 1) & (w_low >= 0)
    ok4 = (h_high <= H - 1) & (w_high <= W - 1)
    return (h_low, w_low, h_high, w_high), (hh, hw, lh, lw), (ok1, ok2, ok3, ok4)


def _gather(img, hi, wi, ok, H, W):
    """img [N, H, W] -> values at (hi, wi) [N, K, Ho, Wo], 0 where !ok"""
    N = img.shape[0]
    hc = np.clip(hi, 0, H - 1)
    wc = np.clip(wi, 0, W - 1)
    n_idx = np.arange(N)[:, None, None, None]
    v = img[n_idx, hc, wc]
    return np.where(ok, v, 0).astype(img.dtype)


def _sample_all(x, h_im, w_im):
    """bilinear samples of every input channel: [N, C, K, Ho, Wo]; zero outside the (-1,H)x(-1,W) window (cuh:177-185)"""
    N, C, H, W = x.shape
    inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
    (h_low, w_low, h_high, w_high), (hh, hw, lh, lw), (ok1, ok2, ok3, ok4) = _corners(h_im, w_im, H, W)
    out = np.zeros((N, C) + h_im.shape[1:], dtype=x.dtype)
    for c in range(C):
        img = x[:, c]
        v1 = _gather(img, h_low, w_low, ok1, H, W)
        v2 = _gather(img, h_low, w_high, ok2, H, W)
        v3 = _gather(img, h_high, w_low, ok3, H, W)
        v4 = _gather(img, h_high, w_high, ok4, H, W)
        val = (hh * hw) * v1 + (hh * lw) * v2 + (lh * hw) * v3 + (lh * lw) * v4
        out[:, c] = np.where(inside, val, 0)
    return out, inside


def mdconv_forward(x, offset, mask, weight, bias, stride=1, pad=1, dil=1):
    """modulated_deform_conv_cuda_forward (modulated_deform_conv_cuda.cu:19-128); numpy arrays in, numpy out [N,Cout,Ho,Wo]"""
    dtype = x.dtype
    N, C, H, W = x.shape
    Cout, Cin, kh, kw = weight.shape
    assert Cin == C, 'groups != 1 is not restated'
    Ho, Wo = _geometry(H, W, kh, kw, stride, pad, dil)
    h_im, w_im = _positions(offset, kh, kw, stride, pad, dil, Ho, Wo, dtype)
    samples, _ = _sample_all(x, h_im, w_im)                     # [N, C, K, Ho, Wo]
    cols = samples * mask[:, None]                              # value * mask (cuh:187)
    wmat = weight.reshape(Cout, C * kh * kw)
    out = np.einsum('ok,nkhw->nohw', wmat, cols.reshape(N, C * kh * kw, Ho, Wo)).astype(dtype)
    return out + bias.reshape(1, Cout, 1, 1).astype(dtype)


def mdconv_backward(x, offset, mask, weight, bias, gout, stride=1, pad=1, dil=1):
    """modulated_deform_conv_cuda_backward (modulated_deform_conv_cuda.cu:131-290) -> (gx, goffset, gmask, gweight, gbias)"""
    dtype = x.dtype
    N, C, H, W = x.shape
    Cout, Cin, kh, kw = weight.shape
    K = kh * kw
    Ho, Wo = _geometry(H, W, kh, kw, stride, pad, dil)
    h_im, w_im = _positions(offset, kh, kw, stride, pad, dil, Ho, Wo, dtype)
    wmat = weight.reshape(Cout, C * K)
    gcol = np.einsum('ok,nohw->nkhw', wmat, gout).reshape(N, C, K, Ho, Wo).astype(dtype)     # columns = W^T * grad_out (:216-221)
    samples, inside = _sample_all(x, h_im, w_im)
    (h_low, w_low, h_high, w_high), (hh, hw, lh, lw), (ok1, ok2, ok3, ok4) = _corners(h_im, w_im, H, W)

    # ---- grad_mask, grad_offset (cuh:254-328): outside the window the position is moved to (-2,-2): no contribution
    gmask = np.where(inside, (gcol * samples).sum(1), 0).astype(dtype)
    goff = np.zeros((N, K, 2, Ho, Wo), dtype=dtype)
    for c in range(C):
        img = x[:, c]
        v1 = _gather(img, h_low, w_low, ok1, H, W)
        v2 = _gather(img, h_low, w_high, ok2, H, W)
        v3 = _gather(img, h_high, w_low, ok3, H, W)
        v4 = _gather(img, h_high, w_high, ok4, H, W)
        # mdmcn_get_coordinate_weight (cuh:86-126); (w_low + 1 - w) == hw, (w - w_low) == lw, ...
        wgt_h = -hw * v1 - lw * v2 + hw * v3 + lw * v4
        wgt_w = -hh * v1 + hh * v2 - lh * v3 + lh * v4
        g = gcol[:, c] * mask
        goff[:, :, 0] += np.where(inside, wgt_h * g, 0)
        goff[:, :, 1] += np.where(inside, wgt_w * g, 0)
    goff = goff.reshape(N, 2 * K, Ho, Wo)

    # ---- grad_input (cuh:196-252): scatter of the bilinear weights; mdmcn_get_gradient_weight is 0 outside the window
    gx = np.zeros_like(x)
    n_idx = np.broadcast_to(np.arange(N)[:, None, None, None], h_im.shape)
    for c in range(C):
        top = np.where(inside, gcol[:, c] * mask, 0)
        for (hi, wi, ok, wt) in ((h_low, w_low, ok1, hh * hw), (h_low, w_high, ok2, hh * lw),
                                 (h_high, w_low, ok3, lh * hw), (h_high, w_high, ok4, lh * lw)):
            sel = ok & inside
            np.add.at(gx[:, c], (n_idx[sel], hi[sel], wi[sel]), (wt * top)[sel].astype(dtype))

    # ---- grad_weight, grad_bias (:262-281)
    cols = (samples * mask[:, None]).reshape(N, C * K, Ho, Wo)
    gw = np.einsum('nohw,nkhw->ok', gout, cols).reshape(weight.shape).astype(dtype)
    gb = gout.sum((0, 2, 3)).astype(dtype)
    return gx, goff, gmask, gw, gb


class MDConvFn(torch.autograd.Function):
    """argument list of ModulatedDeformConvFunction.apply (modulated_deform_conv_func.py:15-56)"""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups, im2col_step):
        assert groups == 1 or input.shape[1] == 1, 'oracle restates groups == 1'
        assert deformable_groups == 1
        ctx.cfg = (int(stride), int(padding), int(dilation))
        ctx.save_for_backward(input, offset, mask, weight, bias)
        out = mdconv_forward(input.detach().numpy(), offset.detach().numpy(), mask.detach().numpy(), weight.detach().numpy(),
                             bias.detach().numpy(), *ctx.cfg)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_output):
        input, offset, mask, weight, bias = ctx.saved_tensors
        g = mdconv_backward(input.detach().numpy(), offset.detach().numpy(), mask.detach().numpy(), weight.detach().numpy(),
                            bias.detach().numpy(), grad_output.contiguous().numpy(), *ctx.cfg)
        return tuple(torch.from_numpy(np.ascontiguousarray(t)) for t in g) + (None,) * 6


# -------------------------------------------------------------------------------------------------------------
# NLSPN (nlspnmodel_adapt.py:189-373), k_f = 3, affinity 'TGASS', conf_prop = True, preserve_input = True
# -------------------------------------------------------------------------------------------------------------
def offset_affinity(offset_aff, confidence, aff_scale_const, legacy=True, conv=MDConvFn.apply):
    """_get_offset_affinity (:255-330).  offset_aff: output of conv_offset_aff(guidance) [B, 24, H, W];
    returns offset [B, 18, H, W] (zero reference offset inserted at tap 4) and affinity [B, 9, H, W]."""
    B, _, H, W = offset_aff.shape
    num, k_f, idx_ref = 8, 3, 4
    o1, o2, aff = torch.chunk(offset_aff, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1).view(B, num, 2, H, W)                       # :262 (a view, not an interleave)
    lst = list(torch.chunk(offset, num, dim=1))
    lst.insert(idx_ref, torch.zeros((B, 1, 2, H, W), dtype=offset.dtype))
    offset = torch.cat(lst, dim=1).view(B, -1, H, W)
    aff = torch.tanh(aff) / (aff_scale_const + 1e-8)                                  # TGASS (:273-274)
    if confidence is not None:
        list_conf = []
        offset_each = torch.chunk(offset, num + 1, dim=1)
        ones = torch.ones((B, 1, H, W), dtype=offset.dtype)
        w_conf = torch.ones((1, 1, 1, 1), dtype=offset.dtype)
        b0 = torch.zeros(1, dtype=offset.dtype)
        for idx_off in range(num + 1):
            ww, hh = idx_off % k_f, idx_off // k_f
            if ww == (k_f - 1) / 2 and hh == (k_f - 1) / 2:
                continue
            off_tmp = offset_each[idx_off].clone().detach()
            if legacy:
                off_tmp[:, 0] = off_tmp[:, 0] + hh - (k_f - 1) / 2
                off_tmp[:, 1] = off_tmp[:, 1] + ww - (k_f - 1) / 2
            list_conf.append(conv(confidence, off_tmp, ones, w_conf, b0, 1, 0, 1, 1, 1, 64))
        aff = aff * torch.cat(list_conf, dim=1).contiguous()
    aff_abs_sum = torch.sum(torch.abs(aff), dim=1, keepdim=True) + 1e-4
    aff_abs_sum = torch.where(aff_abs_sum < 1.0, torch.ones_like(aff_abs_sum), aff_abs_sum)      # in-place masked assignment (:318)
    aff = aff / aff_abs_sum
    aff_ref = 1.0 - torch.sum(aff, dim=1, keepdim=True)
    lst = list(torch.chunk(aff, num, dim=1))
    lst.insert(idx_ref, aff_ref)
    return offset, torch.cat(lst, dim=1)


def propagate(feat_init, offset, aff, feat_fix, prop_time=18, preserve_input=True, conv=MDConvFn.apply):
    """NLSPN.forward propagation loop (:352-373); returns (final feature, list of the per-iteration features)"""
    w = torch.ones((1, 1, 3, 3), dtype=feat_init.dtype)
    b = torch.zeros(1, dtype=feat_init.dtype)
    mask_fix = None
    if preserve_input:
        mask_fix = (torch.sum(feat_fix > 0.0, dim=1, keepdim=True).detach() > 0.0).type_as(feat_fix)
    feat, feats = feat_init, []
    for _ in range(prop_time):
        if preserve_input:
            feat = (1.0 - mask_fix) * feat + mask_fix * feat_fix
        feat = conv(feat, offset, aff, w, b, 1, 1, 1, 1, 1, 64)
        feats.append(feat)
    return feat, feats


def synthetic_prop_inputs(seed, n, h, w, dtype=torch.float32, offset_std=2.5):
    """seeded inputs for the propagation: smooth initial depth, ~5 % sparse depth, random offsets (many samples leave the
    image), affinities normalised as TGASS does"""
    g = torch.Generator().manual_seed(seed)
    yy = torch.arange(h, dtype=dtype).view(1, 1, h, 1) / h
    xx = torch.arange(w, dtype=dtype).view(1, 1, 1, w)
    dense = (5 + 70 * (1 - yy) + 2 * torch.sin(xx / 97)).expand(n, 1, h, w).contiguous()
    feat_init = dense + 0.5 * torch.randn((n, 1, h, w), generator=g, dtype=dtype)
    sparse = dense * (torch.rand((n, 1, h, w), generator=g, dtype=dtype) < 0.05).to(dtype)
    offset_aff = torch.randn((n, 24, h, w), generator=g, dtype=dtype)
    offset_aff[:, :16] *= offset_std
    confidence = torch.sigmoid(torch.randn((n, 1, h, w), generator=g, dtype=dtype))
    return feat_init, sparse, offset_aff, confidence
