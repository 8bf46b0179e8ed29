"""Stage 2 of the source-domain preparation on the NLSPN back-end: training the predictor heads (src/head_main.py:259-275 construction,
:437-480 step; forward `_rgbd_meta_contrast_prepare`, external_src/NLSPN/src/model/nlspnmodel_adapt.py:1014-1060).

One step = both encoders under no_grad with BatchNorm2d in eval mode (src/nlspn_model_adapt.py:360-368 `train_prepare`), the EMA copy
proj_t <- tau proj_t + (1 - tau) proj over proj's six parameters (nlspnmodel_adapt.py:1314-1316), emb = pred(proj(fe6 of the zero image)),
ref = proj_t(fe6 of the frame) (BatchNorm1d of proj / pred in train mode with running-statistics updates, of proj_t in eval mode: the driver
calls convert_syncbn() before the loop, head_main.py:278), loss = mean(2 - 2 cos(emb, ref)) (src/external_model_adapt.py:524-540), gradients
of proj.* and pred.* (only the INPUT of proj is detached, nlspnmodel_adapt.py:1057), torch.optim.Adam.

Everything runs through the C ABI of libptta_b200.so (include/ptta_b200.h): the encoder launches of NlspnEngine, `ptta_gemm_bf16_tc`
(Linear forward + data gradient), `ptta_gemm_tn_bf16_tc` (Linear weight gradient, tcgen05 with MN-major operands), `ptta_nl_bn_stats` /
`ptta_nl_bn_act` / `ptta_nl_bn_backward` (BatchNorm1d + ReLU), `ptta_nl_col_sums` (bias gradients), `ptta_cos_loss_forward` /
`ptta_tta_loss_backward_emb`, `ptta_ema_update`, `ptta_adam_flat_dev`.  No CPU fallback: NlspnEngine raises without the library."""
import copy
from collections import OrderedDict
from ctypes import c_void_p

import torch

from . import _lib
from ._lib import check, ptr

HEAD_TRAINED = tuple('%s.%s.%s' % (m, i, q) for m in ('proj', 'pred') for i in ('0', '1', '3') for q in ('weight', 'bias'))
_EMA_KEYS = ('0.weight', '0.bias', '1.weight', '1.bias', '3.weight', '3.bias')
ACT_RELU = 1
BN_EPS = 1e-5


def fresh_head_state():
    """proj / proj_t / pred as `_prepare_head('head_selfsup_ema')` creates them (nlspnmodel_adapt.py:1338-1343, MLP :1396-1402), drawn from
    torch's global generator in the reference's order -- under the same seed the tensors are the reference's bit for bit
    (oracle/gen_golden_nlspn_prepare.py asserts it)."""
    def mlp(dim, projection_size, hidden_size):
        return torch.nn.Sequential(torch.nn.Linear(dim, hidden_size), torch.nn.BatchNorm1d(hidden_size), torch.nn.ReLU(inplace=True),
                                   torch.nn.Linear(hidden_size, projection_size))
    proj = mlp(512, 1024, 1024)
    proj_t = copy.deepcopy(proj)
    pred = mlp(1024, 1024, 1024)
    sd = OrderedDict()
    for name, m in (('proj', proj), ('proj_t', proj_t), ('pred', pred)):
        for k, v in m.state_dict().items():
            sd['%s.%s' % (name, k)] = v.detach().clone()
    return sd


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class NlspnHeadTrainer:
    """Stage-2 trainer over an NlspnEngine built with syncbn=False (running statistics present)."""

    def __init__(self, engine):
        L = _lib.lib()
        self.eng = e = engine
        self.dev = e.dev
        sd = e.sd
        for k in ('proj.1.running_mean', 'proj_t.1.running_var', 'pred.1.num_batches_tracked'):
            if sd.get(k) is None:
                raise RuntimeError('stage-2 training needs the heads\' running statistics (%s missing)' % k)
        # trained tensors in ONE flat fp32 buffer (+ gradient, Adam moments); the state-dict entries become views
        self.layout, off = {}, 0
        for k in HEAD_TRAINED:
            self.layout[k] = off
            off += sd[k].numel()
        self.flat_p = torch.empty(off, dtype=torch.float32, device=self.dev)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        self.params, self.grads = OrderedDict(), OrderedDict()
        for k in HEAD_TRAINED:
            o, cnt = self.layout[k], sd[k].numel()
            view = self.flat_p[o:o + cnt].view(sd[k].shape)
            view.copy_(sd[k])
            sd[k] = view
            self.params[k] = view
            self.grads[k] = self.flat_g[o:o + cnt].view(sd[k].shape)
        self.adam_hyper = torch.tensor([0.0, 0.9, 0.999, 1e-8, 0.0], dtype=torch.float64, device=self.dev)
        self.adam_step_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._hyper = None
        R = e.R
        self.tn_ws = torch.empty(max(L.ptta_gemm_tn_workspace_bytes(R, 1024, 1024), L.ptta_gemm_tn_workspace_bytes(R, 1024, 512)) // 4,
                                 dtype=torch.float32, device=self.dev)
        if self.tn_ws.numel() == 0:
            raise RuntimeError('ptta_gemm_tn_bf16_tc does not support %d rows x 1024 x {512, 1024}' % R)
        self.loss_ws = torch.zeros(L.ptta_tta_loss_workspace_bytes(e.N, e.H, e.W, R), dtype=torch.uint8, device=self.dev)
        self.ops = {}            # bf16 operand copies of the Linear weights: [out][in] (forward) and [in][out] (data gradient)
        for name in ('proj', 'proj_t', 'pred'):
            self._repack(name)
        self.launches = 0
        self._graph, self._graph_key, self._seen_key = None, None, None

    # ---- helpers ------------------------------------------------------------------------------------------------------------------
    def _repack(self, name):
        sd = self.eng.sd
        for i in ('0', '3'):
            wt = sd['%s.%s.weight' % (name, i)]
            for key, src in (('%s.%s' % (name, i), wt), ('%s.%s.T' % (name, i), wt.t())):
                dst = self.ops.get(key)
                if dst is None:
                    self.ops[key] = src.contiguous().to(torch.bfloat16).contiguous()
                else:
                    dst.copy_(src)

    def set_adam(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        hyper = (float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay))
        if hyper != self._hyper:
            self.adam_hyper.copy_(torch.tensor(hyper, dtype=torch.float64))
            self._hyper = hyper

    def _linear(self, x, name, out_name, n_out):
        e, sd = self.eng, self.eng.sd
        R, k = x.shape
        y = e.buf(out_name, (R, n_out))
        check(_lib.lib().ptta_gemm_bf16_tc(ptr(x), ptr(self.ops[name]), ptr(y), ptr(sd[name + '.bias']), R, n_out, k, _stream()), 'gemm_tc')
        return y

    def _mlp_forward(self, tag, name, x, train_bn):
        """Linear -> BatchNorm1d -> ReLU -> Linear; returns (out, saved) with what the backward needs"""
        e, sd = self.eng, self.eng.sd
        h_raw = self._linear(x, name + '.0', tag + '.h_raw', 1024)
        if train_bn:
            running = (sd[name + '.1.running_mean'], sd[name + '.1.running_var'], sd[name + '.1.num_batches_tracked'])
            st = e.bn_stats(tag + '.bn', h_raw, sd[name + '.1.weight'], sd[name + '.1.bias'], running)
        else:
            st = self._running_state(name + '.1')
        h_act = e.bn_act(h_raw, st, ACT_RELU, tag + '.h_act')
        out = self._linear(h_act, name + '.3', tag + '.out', 1024)
        return out, (x, h_raw, h_act, tag + '.bn')

    def _running_state(self, bn):
        sd = self.eng.sd
        scale = sd[bn + '.weight'] / torch.sqrt(sd[bn + '.running_var'] + BN_EPS)
        return {'scale': scale.contiguous(), 'shift': (sd[bn + '.bias'] - sd[bn + '.running_mean'] * scale).contiguous()}

    def _mlp_backward(self, tag, name, g_out, saved, need_dx):
        """gradients of `name`.{0,1,3}.{weight,bias} into the flat gradient buffer; returns d loss / d input (bf16) if asked"""
        L, e = _lib.lib(), self.eng
        x, h_raw, h_act, bn_key = saved
        R = x.shape[0]
        # Linear .3:  dW = g_out^T h_act,  db = column sums of g_out,  d h_act = g_out W
        check(L.ptta_gemm_tn_bf16_tc(ptr(g_out), ptr(h_act), ptr(self.grads[name + '.3.weight']), ptr(self.tn_ws), R, 1024, 1024, _stream()), 'gemm_tn')
        check(L.ptta_nl_col_sums(ptr(g_out), 1024, R, 1024, ptr(e.partial), ptr(self.grads[name + '.3.bias']), _stream()), 'nl_col_sums')
        d_act = e.buf(tag + '.g_act', (R, 1024))
        check(L.ptta_gemm_bf16_tc(ptr(g_out), ptr(self.ops[name + '.3.T']), ptr(d_act), None, R, 1024, 1024, _stream()), 'gemm_tc')
        # ReLU + BatchNorm1d (train mode)
        st = e.bn_state[bn_key]
        d_raw = e.buf(tag + '.g_raw', (R, 1024))
        check(L.ptta_nl_bn_backward(ptr(d_act), 1024, None, 0, ptr(h_act), ACT_RELU, ptr(h_raw), ptr(st['mean']), ptr(st['rstd']),
                                    ptr(e.sd[name + '.1.weight']), ptr(e.partial), ptr(self.grads[name + '.1.weight']), ptr(self.grads[name + '.1.bias']),
                                    ptr(e.coef), ptr(d_raw), None, R, 1024, _stream()), 'nl_bn_backward')
        # Linear .0
        k = x.shape[1]
        check(L.ptta_gemm_tn_bf16_tc(ptr(d_raw), ptr(x), ptr(self.grads[name + '.0.weight']), ptr(self.tn_ws), R, 1024, k, _stream()), 'gemm_tn')
        check(L.ptta_nl_col_sums(ptr(d_raw), 1024, R, 1024, ptr(e.partial), ptr(self.grads[name + '.0.bias']), _stream()), 'nl_col_sums')
        self.launches += 12
        if not need_dx:
            return None
        dx = e.buf(tag + '.g_in', (R, k))
        check(L.ptta_gemm_bf16_tc(ptr(d_raw), ptr(self.ops[name + '.0.T']), ptr(dx), None, R, k, 1024, _stream()), 'gemm_tc')
        self.launches += 1
        return dx

    # ---- one step --------------------------------------------------------------------------------------------------------------------
    def head_step(self, image_norm, sparse_depth, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_input_depth=None, tau=0.999, graph=False):
        """image_norm: the normalised network input (fp32 NCHW); sparse_depth fp32 [N,1,H,W].  Returns nothing; `read_loss()` reads the loss.
        graph=True: the frame is copied into trainer-owned staging buffers and the step is captured into a CUDA graph the second time it is
        called, then replayed (a fresh tensor may be passed every step; needs a non-default stream)."""
        self.set_adam(lr, betas, eps, weight_decay)
        image_norm, sparse_depth = image_norm.contiguous(), sparse_depth.contiguous()
        if not graph:
            return self._step_body(image_norm, sparse_depth, max_input_depth, tau)
        e = self.eng
        img_in = e.buf('head.stage.image', tuple(image_norm.shape), torch.float32)
        dep_in = e.buf('head.stage.depth', tuple(sparse_depth.shape), torch.float32)
        img_in.copy_(image_norm, non_blocking=True)
        dep_in.copy_(sparse_depth, non_blocking=True)
        key = (max_input_depth, float(tau), e.img_scale is None)
        if self._graph is not None and self._graph_key == key:
            self._graph.replay()
            return
        if self._seen_key != key:
            self._seen_key = key                               # first call: eager (allocates every buffer, sets kernel attributes)
            return self._step_body(img_in, dep_in, max_input_depth, tau)
        g = torch.cuda.CUDAGraph()
        l0 = self.launches + e.launches
        torch.cuda.synchronize(self.dev)
        with torch.cuda.graph(g):
            self._step_body(img_in, dep_in, max_input_depth, tau)
        self.launches_per_step = self.launches + e.launches - l0
        self._graph, self._graph_key = g, key
        g.replay()

    def _step_body(self, image_norm, sparse_depth, max_input_depth, tau):
        L, e = _lib.lib(), self.eng
        sd = e.sd
        R = e.R
        if max_input_depth is not None:                                  # src/external_model_adapt.py:103-108
            d_c = e.buf('head.depth', tuple(sparse_depth.shape), torch.float32)
            check(L.ptta_nl_clamp(ptr(sparse_depth), ptr(d_c), 0.0, float(max_input_depth), d_c.numel(), _stream()), 'nl_clamp')
            sparse_depth = d_c
        # the zero-image encoder and the two trained heads on top of it depend only on the sparse depth: they run on a second stream beside
        # the frame's encoder (fork / join through events, also inside a CUDA-graph capture), as NlspnEngine.forward does for the TTA step
        main, side = torch.cuda.current_stream(), e.side_stream
        e.bn_running = True                                              # frozen encoder: eval-mode BatchNorm2d (running statistics)
        try:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fe6_z = e.encoder('z.', None, sparse_depth)[-1]
                e.bn_running = False
                p_out, p_saved = self._mlp_forward('z.head.proj', 'proj', fe6_z.reshape(R, 512), True)
                emb, q_saved = self._mlp_forward('z.head.pred', 'pred', p_out, True)
                e.bn_running = True
            fe6 = e.encoder('r.', image_norm, sparse_depth)[-1]
        finally:
            e.bn_running = False
        for k in _EMA_KEYS:                                              # nlspnmodel_adapt.py:1055 -> :1314-1316
            t = sd['proj_t.' + k]
            check(L.ptta_ema_update(ptr(t), ptr(sd['proj.' + k]), t.numel(), float(tau), _stream()), 'ema_update')
        self._repack('proj_t')
        ref, _ = self._mlp_forward('head.proj_t', 'proj_t', fe6.reshape(R, 512), False)
        main.wait_stream(side)
        check(L.ptta_cos_loss_forward(ptr(emb), ptr(ref), R, 1024, ptr(self.loss_ws), e.N, e.H, e.W, _stream()), 'cos_loss_forward')
        g_emb = e.buf('head.g_emb', (R, 1024))
        check(L.ptta_tta_loss_backward_emb(ptr(emb), ptr(ref), R, 1024, ptr(self.loss_ws), 1.0, ptr(g_emb), e.N, e.H, e.W, _stream()), 'loss_backward_emb')
        g_p = self._mlp_backward('head.pred', 'pred', g_emb, q_saved, True)
        self._mlp_backward('head.proj', 'proj', g_p, p_saved, False)
        check(L.ptta_adam_flat_dev(ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_m), ptr(self.flat_v), self.flat_p.numel(), ptr(self.adam_hyper),
                                   ptr(self.adam_step_dev), _stream()), 'adam_flat_dev')
        self._repack('proj')
        self._repack('pred')
        self.emb, self.ref = emb, ref
        self.launches += 6 + 3 * 4 + 4

    def read_loss(self):
        return float(self.loss_ws[:4].view(torch.float32).cpu()[0])
