"""Python handle on the native MSG-CHN ProxyTTA engine (include/ptta_b200.h, csrc/engine.cu).

PyTorch is used for device memory and streams only: every tensor handed to the library is a raw
device pointer, every kernel is the library's own."""
import ctypes
from collections import OrderedDict

import torch

from . import _lib
from ._lib import check, ptr, c_void_p

_FLOAT3 = ctypes.c_float * 3


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class MsgChnEngine:
    """One engine = one (N, H, W) shape + one set of weights.

    `state` maps the reference's state-dict keys to fp32 CUDA tensors (int64 for num_batches_tracked); the
    engine reads them in place and updates the BatchNorm buffers in place.  The adapted tensors
    (`adapt_names`) get gradient / Adam-moment buffers that the engine writes."""

    def __init__(self, n, h, w, prepare_mode, state, grads=None, adam_m=None, adam_v=None, options=None, adam_hyper=None):
        self.L = _lib.lib()
        self.n, self.h, self.w, self.prepare_mode = n, h, w, prepare_mode
        handle = c_void_p()
        check(self.L.ptta_msgchn_create(ctypes.byref(handle), n, h, w, prepare_mode.encode()), 'msgchn_create')
        self.handle = handle
        for k, v in (options or {}).items():
            self.set_option(k, v)
        self.device = next(iter(state.values())).device
        nbytes = self.L.ptta_msgchn_workspace_bytes(self.handle)
        self.workspace = torch.empty(nbytes + 512, dtype=torch.uint8, device=self.device)
        off = (-self.workspace.data_ptr()) % 256
        self._ws_off = off
        check(self.L.ptta_msgchn_bind_workspace(self.handle, c_void_p(self.workspace.data_ptr() + off), nbytes, _stream()),
              'bind_workspace')
        self.state = state
        self.grads, self.adam_m, self.adam_v = grads, adam_m, adam_v
        self.adam_hyper = adam_hyper
        self.required_keys = [self.L.ptta_msgchn_key(self.handle, i).decode()
                              for i in range(self.L.ptta_msgchn_num_keys(self.handle))]
        missing = [k for k in self.required_keys if k not in state]
        if missing:
            raise KeyError('state dict lacks %d entries the %s network needs, e.g. %s' % (len(missing), prepare_mode, missing[:3]))
        self.rebind()

    # -- plumbing -------------------------------------------------------------------------------------
    def rebind(self):
        """(Re)send every pointer and rebuild all bf16 operands; call after the state tensors were replaced."""
        for k in self.required_keys:
            t = self.state[k]
            want = torch.int64 if k.endswith('num_batches_tracked') else torch.float32
            if t.dtype != want or not t.is_cuda or not t.is_contiguous():
                raise TypeError('state[%r] must be a contiguous CUDA %s tensor' % (k, want))
            check(self.L.ptta_msgchn_set_tensor(self.handle, k.encode(), ptr(t), t.numel()), 'set_tensor')
        for prefix, d in (('grad/', self.grads), ('adam_m/', self.adam_m), ('adam_v/', self.adam_v)):
            if d:
                for k, t in d.items():
                    check(self.L.ptta_msgchn_set_tensor(self.handle, (prefix + k).encode(), ptr(t), t.numel()), 'set_tensor')
        if self.adam_hyper is not None:
            check(self.L.ptta_msgchn_set_tensor(self.handle, b'adam/hyper', ptr(self.adam_hyper), self.adam_hyper.numel()), 'set_tensor')
        check(self.L.ptta_msgchn_pack_weights(self.handle, _stream()), 'pack_weights')

    def set_option(self, name, value):
        """Kernel-dispatch options of include/ptta_b200.h (`ptta_msgchn_set_option`)."""
        check(self.L.ptta_msgchn_set_option(self.handle, name.encode(), int(value)), 'set_option')

    def set_comm(self, comm):
        """Shared-model mode (sharding.PeerCommunicator): SyncBatchNorm statistics over all ranks and the gradient all-reduce fused with
        the Adam step, through NVLink peer memory inside the engine's own kernels.  None switches it off."""
        check(self.L.ptta_msgchn_set_comm(self.handle, comm.handle if comm is not None else None), 'set_comm')
        self._comm = comm               # keep the mapping alive as long as the engine uses it

    def close(self):
        if getattr(self, 'handle', None):
            torch.cuda.synchronize(self.device)
            self.L.ptta_msgchn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tensor_names(self):
        return [self.L.ptta_msgchn_tensor_name(self.handle, i).decode() for i in range(self.L.ptta_msgchn_num_tensors(self.handle))]

    def tensor(self, name):
        """Zero-copy torch view of an engine-owned tensor (NHWC for maps)."""
        p, dt = c_void_p(), ctypes.c_int()
        dims = (ctypes.c_longlong * 4)()
        check(self.L.ptta_msgchn_get_tensor(self.handle, name.encode(), ctypes.byref(p), ctypes.byref(dt), dims), 'get_tensor')
        dtype = torch.float32 if dt.value == 0 else torch.bfloat16
        shape = [int(d) for d in dims]
        numel = 1
        for d in shape:
            numel *= d
        off = p.value - self.workspace.data_ptr()
        nbytes = numel * (4 if dt.value == 0 else 2)
        return self.workspace[off:off + nbytes].view(dtype).view(shape)

    # -- the step ----------------------------------------------------------------------------------------
    @staticmethod
    def _f3(v):
        return _FLOAT3(*[float(x) for x in v])

    def forward(self, image, sparse_depth, max_input_depth, training, img_scale=(1.0, 1.0, 1.0), img_shift=(0.0, 0.0, 0.0)):
        """training: False / 0 eval, True / 1 train with the zero-image branch and the proxy heads, 2 train without them (stage 1)"""
        self._check_inputs(image, sparse_depth)
        cap = float(max_input_depth) if max_input_depth is not None else -1.0
        check(self.L.ptta_msgchn_forward(self.handle, ptr(image), self._f3(img_scale), self._f3(img_shift), ptr(sparse_depth), cap,
                                         int(training), _stream()), 'forward')

    # -- source-domain preparation (include/ptta_b200.h, "source-domain preparation steps") -----------------
    def l2_loss(self, ground_truth, max_predict_depth):
        self._check_map(ground_truth)
        check(self.L.ptta_msgchn_l2_loss(self.handle, ptr(ground_truth), float(max_predict_depth), _stream()), 'l2_loss')

    def l2_loss_backward(self, grad_scale=1.0):
        check(self.L.ptta_msgchn_l2_loss_backward(self.handle, grad_scale, _stream()), 'l2_loss_backward')

    def init_step(self, image_raw, sparse_depth, ground_truth, max_input_depth, max_predict_depth, img_scale, img_shift, graph=False):
        self._check_inputs(image_raw, sparse_depth)
        self._check_map(ground_truth)
        cap = float(max_input_depth) if max_input_depth is not None else -1.0
        fn = self.L.ptta_msgchn_init_step_graph if graph else self.L.ptta_msgchn_init_step
        check(fn(self.handle, ptr(image_raw), self._f3(img_scale), self._f3(img_shift), ptr(sparse_depth),
                                           ptr(ground_truth), cap, float(max_predict_depth), _stream()), 'init_step')

    def cos_loss(self):
        check(self.L.ptta_msgchn_cos_loss(self.handle, _stream()), 'cos_loss')

    def cos_loss_backward(self, grad_scale=1.0):
        check(self.L.ptta_msgchn_cos_loss_backward(self.handle, grad_scale, _stream()), 'cos_loss_backward')

    def ema_update_head(self, tau=0.999):
        check(self.L.ptta_msgchn_ema_update_head(self.handle, tau, _stream()), 'ema_update_head')

    def head_backward(self):
        check(self.L.ptta_msgchn_head_backward(self.handle, _stream()), 'head_backward')

    def head_step(self, image_raw, sparse_depth, max_input_depth, img_scale, img_shift, graph=False):
        self._check_inputs(image_raw, sparse_depth)
        cap = float(max_input_depth) if max_input_depth is not None else -1.0
        fn = self.L.ptta_msgchn_head_step_graph if graph else self.L.ptta_msgchn_head_step
        check(fn(self.handle, ptr(image_raw), self._f3(img_scale), self._f3(img_shift), ptr(sparse_depth), cap,
                                           _stream()), 'head_step')

    def _check_map(self, t):
        if tuple(t.shape) != (self.n, 1, self.h, self.w) or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise TypeError('expected a contiguous fp32 CUDA tensor of shape %s' % ((self.n, 1, self.h, self.w),))

    def loss(self, image_raw, sparse_depth, validity, max_input_depth, w_sd, w_sm, w_cos):
        cap = float(max_input_depth) if max_input_depth is not None else -1.0
        check(self.L.ptta_msgchn_loss(self.handle, ptr(image_raw), ptr(sparse_depth), ptr(validity), cap, w_sd, w_sm, w_cos, _stream()),
              'loss')

    def read_losses(self):
        out = (ctypes.c_float * 5)()
        check(self.L.ptta_msgchn_read_losses(self.handle, out, _stream()), 'read_losses')
        return {'loss': out[0], 'loss_sparse_depth': out[1], 'loss_smooth': out[2], 'loss_cos': out[3], 'w_cos_eff': out[4]}

    def backward(self, grad_scale=1.0):
        check(self.L.ptta_msgchn_backward(self.handle, grad_scale, _stream()), 'backward')

    def loss_backward(self, grad_scale=1.0):
        check(self.L.ptta_msgchn_loss_backward(self.handle, grad_scale, _stream()), 'loss_backward')

    def network_backward(self):
        check(self.L.ptta_msgchn_network_backward(self.handle, _stream()), 'network_backward')

    def set_adam(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, step_count=-1):
        check(self.L.ptta_msgchn_set_adam(self.handle, lr, betas[0], betas[1], eps, weight_decay, step_count, _stream()), 'set_adam')

    def adam_step(self):
        check(self.L.ptta_msgchn_adam_step(self.handle, _stream()), 'adam_step')

    def pack_adapted(self):
        check(self.L.ptta_msgchn_pack_adapted(self.handle, _stream()), 'pack_adapted')

    def tta_step(self, image_raw, sparse_depth, max_input_depth, w_sd, w_sm, w_cos, img_scale, img_shift, graph=False):
        self._check_inputs(image_raw, sparse_depth)
        cap = float(max_input_depth) if max_input_depth is not None else -1.0
        fn = self.L.ptta_msgchn_tta_step_graph if graph else self.L.ptta_msgchn_tta_step
        check(fn(self.handle, ptr(image_raw), self._f3(img_scale), self._f3(img_shift), ptr(sparse_depth), cap, w_sd, w_sm, w_cos,
                 _stream()), 'tta_step')

    def launch_count(self):
        return int(self.L.ptta_msgchn_launch_count(self.handle))

    def _check_inputs(self, image, sparse_depth):
        if tuple(image.shape) != (self.n, 3, self.h, self.w) or tuple(sparse_depth.shape) != (self.n, 1, self.h, self.w):
            raise ValueError('engine built for N=%d %dx%d, got image %s / sparse depth %s' % (
                self.n, self.h, self.w, tuple(image.shape), tuple(sparse_depth.shape)))
        for t in (image, sparse_depth):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise TypeError('inputs must be contiguous fp32 CUDA tensors')
