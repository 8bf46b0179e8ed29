"""Host-side mirror of the NLSPN propagation path on top of libptta_b200.so.

  * `ModulatedDeformConvFunction` -- drop-in for external_src/NLSPN/src/model/modulated_deform_conv_func.py:15-56 (the autograd
    Function over the reference's DCN pybind module): same argument list, same returned gradients, same error behaviour
    (RuntimeError for non-contiguous / non-CUDA tensors, cuda/modulated_deform_conv_cuda.cu:39-73).
  * `offset_affinity` -- NLSPN._get_offset_affinity (nlspnmodel_adapt.py:255-330) as ONE fused kernel per direction; input is the
    output of `conv_offset_aff(guidance)`.
  * `propagate` -- the prop_time-step loop of NLSPN.forward (:352-373) with the input-preserving blend, fused (no columns
    buffer, no per-step autograd nodes), differentiable with respect to the initial depth, the offsets and the affinities.
  * `NLSPNPropagation` -- nn.Module with NLSPN's forward signature for everything after `conv_offset_aff`.

No CPU or PyTorch fallback: CPU tensors raise."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _need_cuda_f32(t, name, contiguous=True):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor' % name)
    if t.dtype != torch.float32:
        raise RuntimeError('%s must be float32 (got %s)' % (name, t.dtype))
    if contiguous and not t.is_contiguous():
        raise RuntimeError('%s tensor has to be contiguous' % name)


class ModulatedDeformConvFunction(Function):
    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups, im2col_step):
        for t, nm in ((input, 'input'), (weight, 'weight'), (bias, 'bias'), (offset, 'offset'), (mask, 'mask')):
            _need_cuda_f32(t, nm)
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        ctx.groups, ctx.deformable_groups = groups, deformable_groups
        n, c, h, w = input.shape
        cout, cin_k, kh, kw = weight.shape
        if c != cin_k * groups:
            raise RuntimeError('Input shape and kernel channels wont match: (%d vs %d).' % (c, cin_k * groups))
        if ctx.stride[0] != ctx.stride[1] or ctx.padding[0] != ctx.padding[1] or ctx.dilation[0] != ctx.dilation[1]:
            raise RuntimeError('anisotropic stride / padding / dilation is not implemented')
        if ctx.stride[0] != 1 or ctx.dilation[0] != 1:
            raise NotImplementedError('ModulatedDeformConvFunction: stride / dilation other than 1 are not on the NLSPN path (got %d / %d)' % (
                ctx.stride[0], ctx.dilation[0]))
        ho, wo = h + 2 * ctx.padding[0] - (kh - 1), w + 2 * ctx.padding[1] - (kw - 1)
        if tuple(offset.shape) != (n, 2 * kh * kw * deformable_groups, ho, wo) or tuple(mask.shape) != (n, kh * kw * deformable_groups, ho, wo):
            raise RuntimeError('offset / mask shape does not match the output size %dx%d' % (ho, wo))
        out = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=input.device)
        check(_lib.lib().ptta_mdconv_forward(ptr(input), ptr(weight), ptr(bias), ptr(offset), ptr(mask), ptr(out), n, c, h, w, cout, kh, kw,
                                             ctx.stride[0], ctx.padding[0], ctx.dilation[0], groups, deformable_groups, _stream()), 'mdconv_forward')
        ctx.save_for_backward(input, offset, mask, weight, bias)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask, weight, bias = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        n, c, h, w = input.shape
        cout, _, kh, kw = weight.shape
        gi, go, gm = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
        gw, gb = torch.empty_like(weight), torch.empty_like(bias)
        check(_lib.lib().ptta_mdconv_backward(ptr(input), ptr(weight), ptr(offset), ptr(mask), ptr(grad_output), ptr(gi), ptr(go), ptr(gm), ptr(gw),
                                              ptr(gb), n, c, h, w, cout, kh, kw, ctx.stride[0], ctx.padding[0], ctx.dilation[0], ctx.groups,
                                              ctx.deformable_groups, _stream()), 'mdconv_backward')
        return gi, go, gm, gw, gb, None, None, None, None, None, None


class _OffsetAffinity(Function):
    @staticmethod
    def forward(ctx, offset_aff, confidence, aff_scale_const, legacy):
        _need_cuda_f32(offset_aff, 'offset_aff')
        if confidence is not None:
            _need_cuda_f32(confidence, 'confidence')
        n, ch, h, w = offset_aff.shape
        if ch != 24:
            raise RuntimeError('offset_aff must have 3 * 8 channels (k_f = 3), got %d' % ch)
        offset = torch.empty((n, 18, h, w), dtype=torch.float32, device=offset_aff.device)
        aff = torch.empty((n, 9, h, w), dtype=torch.float32, device=offset_aff.device)
        check(_lib.lib().ptta_nlspn_offset_affinity_forward(ptr(offset_aff), ptr(confidence), float(aff_scale_const), int(bool(legacy)), ptr(offset),
                                                            ptr(aff), n, h, w, _stream()), 'nlspn_offset_affinity_forward')
        ctx.save_for_backward(offset_aff, confidence)
        ctx.cfg = (float(aff_scale_const), int(bool(legacy)))
        return offset, aff

    @staticmethod
    @once_differentiable
    def backward(ctx, g_offset, g_aff):
        offset_aff, confidence = ctx.saved_tensors
        n, _, h, w = offset_aff.shape
        g_oa = torch.empty_like(offset_aff)
        g_conf = torch.empty_like(confidence) if confidence is not None else None
        check(_lib.lib().ptta_nlspn_offset_affinity_backward(ptr(offset_aff), ptr(confidence), ctx.cfg[0], ctx.cfg[1], ptr(g_offset.contiguous()),
                                                             ptr(g_aff.contiguous()), ptr(g_oa), ptr(g_conf), n, h, w, _stream()),
              'nlspn_offset_affinity_backward')
        return g_oa, g_conf, None, None


def offset_affinity(offset_aff, confidence, aff_scale_const, legacy=True):
    """(offset [N,18,H,W], aff [N,9,H,W]) from conv_offset_aff's output; the gradient with respect to `aff_scale_const` is not
    produced (it is frozen in every adapt mode of the TTA driver, src/nlspn_model_adapt.py:306-333)"""
    return _OffsetAffinity.apply(offset_aff, confidence, aff_scale_const, legacy)


class _Propagate(Function):
    @staticmethod
    def forward(ctx, feat_init, offset, aff, feat_fix, prop_time, want_list):
        for t, nm in ((feat_init, 'feat_init'), (offset, 'offset'), (aff, 'aff')):
            _need_cuda_f32(t, nm)
        if feat_fix is not None:
            _need_cuda_f32(feat_fix, 'feat_fix')
        n, c, h, w = feat_init.shape
        if c != 1:
            raise RuntimeError('only tested with ch_f == 1 but %d' % c)          # the reference's own assertion (nlspnmodel_adapt.py:199)
        if tuple(offset.shape) != (n, 18, h, w) or tuple(aff.shape) != (n, 9, h, w):
            raise RuntimeError('offset / affinity shape mismatch')
        out = torch.empty_like(feat_init)
        saved = torch.empty((prop_time, n, h, w), dtype=torch.float32, device=feat_init.device)
        lst = torch.empty((prop_time, n, 1, h, w), dtype=torch.float32, device=feat_init.device) if want_list else None
        check(_lib.lib().ptta_nlspn_propagate_forward(ptr(feat_init), ptr(offset), ptr(aff), ptr(feat_fix), ptr(out), ptr(saved), ptr(lst), n, h, w,
                                                      prop_time, _stream()), 'nlspn_propagate_forward')
        ctx.save_for_backward(offset, aff, feat_fix, saved)
        ctx.prop_time = prop_time
        if want_list:
            ctx.mark_non_differentiable(lst)
            return out, lst
        return out, None

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out, _g_list):
        offset, aff, feat_fix, saved = ctx.saved_tensors
        n, _, h, w = offset.shape
        g_init = torch.empty((n, 1, h, w), dtype=torch.float32, device=offset.device)
        g_off, g_aff = torch.empty_like(offset), torch.empty_like(aff)
        scratch = torch.empty((2, n, h, w), dtype=torch.float32, device=offset.device)
        check(_lib.lib().ptta_nlspn_propagate_backward(ptr(g_out.contiguous()), ptr(offset), ptr(aff), ptr(feat_fix), ptr(saved), ptr(g_init),
                                                       ptr(g_off), ptr(g_aff), ptr(scratch), n, h, w, ctx.prop_time, _stream()),
              'nlspn_propagate_backward')
        return g_init, g_off, g_aff, None, None, None


def propagate(feat_init, offset, aff, feat_fix=None, prop_time=18, return_list=False):
    """final feature (and, with return_list, the per-step features [prop_time, N, 1, H, W], not differentiable)"""
    out, lst = _Propagate.apply(feat_init, offset, aff, feat_fix, prop_time, return_list)
    return (out, lst) if return_list else out


class NLSPNPropagation(torch.nn.Module):
    """NLSPN (nlspnmodel_adapt.py:189-373) after `conv_offset_aff`: forward(feat_init, offset_aff, confidence, feat_fix) ->
    (feat_result, list_feat, offset, aff, aff_scale_const) exactly as NLSPN.forward returns them."""

    def __init__(self, prop_time=18, affinity_gamma=0.5, conf_prop=True, legacy=True, preserve_input=True):
        super().__init__()
        self.prop_time, self.conf_prop, self.legacy, self.preserve_input = prop_time, conf_prop, legacy, preserve_input
        self.aff_scale_const = torch.nn.Parameter(affinity_gamma * 8 * torch.ones(1), requires_grad=False)

    def forward(self, feat_init, offset_aff, confidence=None, feat_fix=None):
        if self.conf_prop and confidence is None:
            raise AssertionError('conf_prop needs a confidence map')
        offset, aff = offset_affinity(offset_aff, confidence if self.conf_prop else None, float(self.aff_scale_const), self.legacy)
        out, lst = propagate(feat_init, offset, aff, feat_fix if self.preserve_input else None, self.prop_time, return_list=True)
        return out, list(lst.unbind(0)), offset, aff, self.aff_scale_const.data
