"""torch-tensor wrappers over the general-channel tcgen05 convolution family (csrc/conv_gen.cuh, `ptta_convg_*`):
the conv / transposed-conv layers of the NLSPN network (external_src/NLSPN/src/model/nlspnmodel_adapt.py:384-448) and their
data gradients.  Feature maps are NHWC bf16 [N,H,W,C] with C a multiple of 64.  No PyTorch fallback."""
import torch

from . import _lib
from ._lib import check, ptr, c_void_p

KINDS = {'s1': 0, 's2': 1, 't2': 2, 'p1s2': 3}
FWD, DGRAD = 0, 1


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _pad64(c):
    return (c + 63) // 64 * 64


class ConvG:
    """One layer in one role: owns the packed bf16 weights; `__call__` launches the kernel.

    kind: 's1' Conv2d 3x3 s1 p1 | 's2' Conv2d 3x3 s2 p1 | 't2' ConvTranspose2d 3x3 s2 p1 op1 | 'p1s2' Conv2d 1x1 s2.
    role: FWD or DGRAD.  cin = (cin0, cin1) stored channels of the one or two (concatenated) inputs, cout stored output channels.
    weight: the reference's fp32 parameter; weight_short: the 1x1/s2 shortcut weight whose data gradient is folded in."""

    def __init__(self, kind, role, weight, cin, cout, weight_short=None, bias=None, ident_from=-1):
        self.kind, self.role = KINDS[kind], role
        self.cin0, self.cin1 = (cin, 0) if isinstance(cin, int) else cin
        self.cout = cout
        self.has_short = 0 if weight_short is None else 1
        self.ident_from = ident_from
        self.weight, self.weight_short = weight, weight_short
        if kind == 't2':
            self.cin_w, self.cout_w = weight.shape[0], weight.shape[1]
        else:
            self.cout_w, self.cin_w = weight.shape[0], weight.shape[1]
        n = _lib.lib().ptta_convg_packed_elems(self.kind, role, self.cin0, self.cin1, cout, self.has_short)
        if n < 0:
            raise RuntimeError('convg: ' + _lib.last_error())
        self.packed = torch.empty(n, dtype=torch.bfloat16, device=weight.device)
        self.bias = None
        if bias is not None:
            self.bias = torch.zeros(cout if role == FWD else self.cin0, dtype=torch.float32, device=weight.device)
            self.bias[:bias.numel()] = bias
        self.repack()

    def repack(self):
        w = self.weight.detach()
        ws = None if self.weight_short is None else self.weight_short.detach()
        assert w.is_contiguous() and w.dtype == torch.float32 and (ws is None or (ws.is_contiguous() and ws.dtype == torch.float32))
        check(_lib.lib().ptta_convg_pack(self.kind, self.role, ptr(w), ptr(ws), self.cin_w, self.cout_w, self.cin0, self.cin1, self.cout,
                                         self.has_short, self.ident_from, ptr(self.packed), _stream()), 'convg_pack')

    def out_shape(self, n, h, w):
        """h, w: LAYER INPUT size"""
        if self.role == FWD:
            if self.kind == 0:
                return (n, h, w, self.cout)
            if self.kind == 2:
                return (n, 2 * h, 2 * w, self.cout)
            return (n, h // 2, w // 2, self.cout)
        return (n, h, w, self.cin0)

    def __call__(self, x0, x1=None, out=None, hw=None):
        """FWD: x0 (and x1) are the layer input(s).  DGRAD: x0 = gradient of the layer output (x1 = of the shortcut output) and
        `hw` = the layer INPUT size."""
        n = x0.shape[0]
        if self.role == FWD:
            h, w = x0.shape[1], x0.shape[2]
        else:
            h, w = hw
        for t in (x0, x1):
            if t is not None and (t.dtype != torch.bfloat16 or not t.is_contiguous() or not t.is_cuda):
                raise TypeError('convg operands must be contiguous CUDA bf16 NHWC tensors')
        if out is None:
            out = torch.empty(self.out_shape(n, h, w), dtype=torch.bfloat16, device=x0.device)
        check(_lib.lib().ptta_convg_run(self.kind, self.role, ptr(x0), ptr(x1), ptr(self.packed), ptr(self.bias), ptr(out), n, h, w,
                                        self.cin0, self.cin1, self.cout, self.has_short, _stream()), 'convg_run')
        return out
