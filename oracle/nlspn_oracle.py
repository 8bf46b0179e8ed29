"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of ProxyTTA's per-frame adaptation step for the
NLSPN back-end (SURVEY.md section 8 row a18 on top of a19-a21).  It is the checker for the CUDA path, never the thing
shipped or measured: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import it.

Parity pinning: the reference holds no golden vectors; this file is pinned against *outputs of the reference itself*
(`oracle/gen_golden_nlspn_net.py` instantiates the reference's own `ExternalModel_Adapt('nlspn')`, loads the seeded
checkpoint produced by `make_synthetic_checkpoint` below, runs the driver's forward / compute_loss / backward / Adam
lines and commits losses, outputs, gradients and adapted tensors under tests/golden/nlspn_net_*.pt).
tests/test_nlspn_net_oracle.py checks this file against those fixtures.  The reference's DCN CUDA extension has no CPU
path; in the fixture run that one import is resolved to oracle.nlspn_prop_oracle.MDConvFn (pinned separately against
torchvision's DCNv2 in fp64 and against the extension itself on the GPU box).

Functional over a flat state dict whose keys and shapes are those of the reference's `NLSPNModel_Adapt.state_dict()`
after `_prepare_head('meta_selfsup_seq_1layer_ema')` (oracle/nlspn_state_manifest.json, written by the generator).

Reference files restated here (paths relative to the reference root):
  M  = external_src/NLSPN/src/model/nlspnmodel_adapt.py
  C  = external_src/NLSPN/src/model/common.py
  W  = src/nlspn_model_adapt.py
  E  = src/external_model_adapt.py
  T  = src/tta_main.py
  torchvision.models.resnet34 (BasicBlock, layers [3, 4, 6, 3]) -- third-party, pinned torchvision==0.10.1 by the
  reference's README; restated from its published definition (conv3x3-BN-ReLU-conv3x3-BN (+1x1/s2 conv-BN shortcut)-add-ReLU).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import msgchn_oracle as O
from . import nlspn_prop_oracle as P

PREPARE_MODE = 'meta_selfsup_seq_1layer_ema'
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
RESNET34_LAYERS = (('conv2', 64, 64, 3, 1), ('conv3', 64, 128, 4, 2), ('conv4', 128, 256, 6, 2), ('conv5', 256, 512, 3, 2))
BN_EPS = 1e-5


class Precision(O.Precision):
    def wgt(self, w):          # every NLSPN conv / linear weight is a bf16 tensor-core operand on the native path
        return w.to(self.dtype).to(torch.float32) if self.dtype is not None else w


FP32 = O.FP32


# ----------------------------------------------------------------------------------------------------------------
# synthetic checkpoint / frames: input generators shared with bench.py, re-exported
# ----------------------------------------------------------------------------------------------------------------
from tta_depth_completion_b200.synthetic import (make_nlspn_checkpoint as make_synthetic_checkpoint, synthetic_frame,  # noqa: E402,F401
                                                 normalize_image_imagenet as normalize_image)


# ----------------------------------------------------------------------------------------------------------------
# network pieces
# ----------------------------------------------------------------------------------------------------------------
_BN_RUNNING = [False]


class running_batchnorm:
    """stage 2 (src/head_main.py): `train_prepare()` (W:360-368) puts every BatchNorm outside proj / pred into eval mode and nothing has
    removed the running statistics there (adapt_parameters is not called), so the frozen encoder normalises with them"""
    def __enter__(self):
        _BN_RUNNING[0] = True

    def __exit__(self, *a):
        _BN_RUNNING[0] = False


def _bn2d(sd, name, x):
    """Every BatchNorm2d after adapt_parameters('meta_bn') (W:322-337): batch statistics in train AND eval, no running
    statistics (track_running_stats False, buffers None)."""
    if _BN_RUNNING[0]:
        return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'], sd[name + '.weight'], sd[name + '.bias'], False, 0.1, BN_EPS)
    return F.batch_norm(x, None, None, sd[name + '.weight'], sd[name + '.bias'], True, 0.1, BN_EPS)


def _cbr(sd, name, x, pr, stride=1, bn=True, act='leaky'):
    # C:45-60 conv_bn_relu: Conv2d(bias = not bn) [+ BN] [+ LeakyReLU(0.2)]
    y = F.conv2d(pr.act(x), pr.wgt(sd[name + '.0.weight']), sd.get(name + '.0.bias'), stride=stride, padding=1)
    if bn:
        y = _bn2d(sd, name + '.1', y)
    if act == 'leaky':
        y = F.leaky_relu(y, 0.2)
    return y


def _ctbr(sd, name, x, pr):
    # C:63-80 convt_bn_relu: ConvTranspose2d(k3, s2, p1, op1, no bias) + BN + LeakyReLU(0.2)
    y = F.conv_transpose2d(pr.act(x), pr.wgt(sd[name + '.0.weight']), None, stride=2, padding=1, output_padding=1)
    return F.leaky_relu(_bn2d(sd, name + '.1', y), 0.2)


def _basic_block(sd, p, x, stride, pr):
    # torchvision BasicBlock.forward
    out = F.conv2d(pr.act(x), pr.wgt(sd[p + '.conv1.weight']), None, stride=stride, padding=1)
    out = F.relu(_bn2d(sd, p + '.bn1', out))
    out = F.conv2d(pr.act(out), pr.wgt(sd[p + '.conv2.weight']), None, stride=1, padding=1)
    out = _bn2d(sd, p + '.bn2', out)
    if (p + '.downsample.0.weight') in sd:
        idt = F.conv2d(pr.act(x), pr.wgt(sd[p + '.downsample.0.weight']), None, stride=stride)
        idt = _bn2d(sd, p + '.downsample.1', idt)
    else:
        idt = x
    return F.relu(out + idt)


def _res_layer(sd, name, x, blocks, stride, pr):
    for b in range(blocks):
        x = _basic_block(sd, '%s.%d' % (name, b), x, stride if b == 0 else 1, pr)
    return x


def _concat(fd, fe):
    # M:473-490: crop the decoder feature if it is larger than the encoder one
    hd, wd = fd.shape[-2:]
    he, we = fe.shape[-2:]
    if hd > he:
        fd = fd[:, :, :he, :]
    if wd > we:
        fd = fd[:, :, :, :we]
    return torch.cat((fd, fe), dim=1)


def encoder(sd, image, sparse_depth, pr=FP32):
    """M:866-880: fe1 .. fe6"""
    x = _cbr(sd, 'conv1_rgb', image, pr, bn=False)
    fe1_rgb = F.conv2d(pr.act(x), pr.wgt(sd['conv1_rgb_meta.weight']), sd['conv1_rgb_meta.bias'], padding=1)   # '1layer' meta conv
    fe1_dep = _cbr(sd, 'conv1_dep', sparse_depth, pr, bn=False)                                               # conv1_dep_meta = Identity
    fe = [torch.cat((fe1_rgb, fe1_dep), dim=1)]
    for name, cin, cout, blocks, stride in RESNET34_LAYERS:
        fe.append(_res_layer(sd, name, fe[-1], blocks, stride, pr))
    fe.append(_cbr(sd, 'conv6', fe[-1], pr, stride=2))
    return fe


def _mlp(sd, name, x, training, pr):
    """MLP head (M:1396-1402) in the state the driver leaves it in: convert_syncbn() (T:327) makes its BatchNorm1d a SyncBatchNorm, so
    adapt_parameters('meta_bn') (W:328-337) sets its running statistics to None -- batch statistics in train and eval, nothing to
    update -- and returns its affine pair among the adapted tensors."""
    h = pr.act(F.linear(x, pr.wgt(sd[name + '.0.weight']), sd[name + '.0.bias']))
    h = pr.act(F.relu(F.batch_norm(h, None, None, sd[name + '.1.weight'], sd[name + '.1.bias'], True, 0.1, 1e-5)))
    return pr.act(F.linear(h, pr.wgt(sd[name + '.3.weight']), sd[name + '.3.bias']))


def network_forward(sd, image, sparse_depth, training, pr=FP32, legacy=True, prop_time=18, trace=None):
    """NLSPNModel_Adapt._rgbd_meta_contrast with mode = [adapt, seq, reverse, ema] (M:850-944)."""
    fe1, fe2, fe3, fe4, fe5, fe6 = encoder(sd, image, sparse_depth, pr)
    fd5 = _ctbr(sd, 'dec5', fe6, pr)
    fd4 = _ctbr(sd, 'dec4', _concat(fd5, fe5), pr)
    fd3 = _ctbr(sd, 'dec3', _concat(fd4, fe4), pr)
    fd2 = _ctbr(sd, 'dec2', _concat(fd3, fe3), pr)
    f21 = _concat(fd2, fe2)
    id_fd1 = _cbr(sd, 'id_dec1', f21, pr)
    pred_init = _cbr(sd, 'id_dec0', _concat(id_fd1, fe1), pr, bn=False)           # conv + LeakyReLU (M:430-431)
    gd_fd1 = _cbr(sd, 'gd_dec1', f21, pr)
    guide = _cbr(sd, 'gd_dec0', _concat(gd_fd1, fe1), pr, bn=False, act=None)
    cf_fd1 = _cbr(sd, 'cf_dec1', f21, pr)
    confidence = torch.sigmoid(F.conv2d(pr.act(_concat(cf_fd1, fe1)), pr.wgt(sd['cf_dec0.0.weight']), sd['cf_dec0.0.bias'], padding=1))
    # prop_layer (M:340-373)
    offset_aff = F.conv2d(guide, sd['prop_layer.conv_offset_aff.weight'], sd['prop_layer.conv_offset_aff.bias'], padding=1)
    offset, aff = P.offset_affinity(offset_aff, confidence, sd['prop_layer.aff_scale_const'], legacy=legacy)
    y, _ = P.propagate(pred_init, offset, aff, sparse_depth, prop_time=prop_time, preserve_input=True)
    output = torch.clamp(y, min=0)                                                 # M:901
    if trace is not None:
        trace.update(fe1=fe1, fe2=fe2, fe3=fe3, fe4=fe4, fe5=fe5, fe6=fe6, fd5=fd5, fd4=fd4, fd3=fd3, fd2=fd2, pred_init=pred_init,
                     guide=guide, confidence=confidence, offset=offset, aff=aff, y=y)
    if not training:
        return output
    with torch.no_grad():                                                          # M:905-914
        fe6_z = encoder(sd, torch.zeros_like(image), sparse_depth, pr)[-1]
    # M:935-936 ('ema', 'reverse', 'adapt'): rows are pixels of the /16 map in NHWC order
    z_zero = fe6_z.permute(0, 2, 3, 1).reshape(-1, 512).detach()
    z_real = fe6.permute(0, 2, 3, 1).reshape(-1, 512)
    emb = _mlp(sd, 'pred', _mlp(sd, 'proj', z_zero, training, pr), training, pr)
    ref = _mlp(sd, 'proj_t', z_real, training, pr)
    if trace is not None:
        trace.update(fe6_zero=fe6_z, emb=emb, ref=ref)
    return output, emb, ref


def model_forward(sd, image, sparse_depth, training, max_input_depth, pr=FP32, trace=None):
    """ExternalModel_Adapt.forward (E:103-108: clamp) -> NLSPNModel_Adapt.forward (W:88-128; the eval-time CPU
    `inpainting` of W:124-127 is not part of the adaptation step)."""
    if max_input_depth is not None:
        sparse_depth = torch.clamp(sparse_depth, 0, max_input_depth)
    return network_forward(sd, image, sparse_depth, training, pr, trace=trace)


# ----------------------------------------------------------------------------------------------------------------
# adapted parameters (W:322-337) and one TTA step (T:583-633)
# ----------------------------------------------------------------------------------------------------------------
def adapt_parameter_names(sd, mode='meta_bn'):
    """'meta_bn' after convert_syncbn() (the driver's sequence, T:327-339): every parameter whose name contains 'meta', then weight and
    bias of every (Sync)BatchNorm in module order -- the BatchNorm2d layers of the network AND the BatchNorm1d layers of the three heads
    (the `'pred' not in np and 'proj' not in np` test of W:335 looks at the LOCAL parameter name, 'weight' / 'bias', and excludes nothing)."""
    if mode != 'meta_bn':
        raise NotImplementedError(mode)
    names = [k for k in sd if 'meta' in k and k.rsplit('.', 1)[-1] in ('weight', 'bias')]
    for k in sd:
        if k.endswith('.running_mean'):
            base = k[:-len('.running_mean')]
            names += [base + '.weight', base + '.bias']
    return names


def tta_step(sd, state, image, sparse_depth, *, lr=3e-4, w_sd=1.0, w_sm=1.0, w_cos=0.1, max_input_depth=80.0,
             pr=FP32, return_grads=False, trace=None):
    """`image` is the raw [0,255] image; the network sees the ImageNet-normalised one, the smoothness loss the raw one.
    Mutates `sd` (adapted tensors, BatchNorm1d running statistics of the heads) and `state`."""
    names = list(state.m.keys())
    v = O.validity_map(sparse_depth)
    d_f, v_f = O.remove_outliers(sparse_depth, v)
    work = dict(sd)
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    out, emb, ref = model_forward(work, normalize_image(image), d_f, True, max_input_depth, pr, trace=trace)
    loss, info = O.adapt_loss(image, out, d_f, v_f, emb, ref, w_sd, w_sm, w_cos, max_input_depth)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    O.adam_update(sd, grads, state, lr)
    res = {'validity': v_f, 'sparse_depth': d_f, 'output_depth': out.detach(), 'loss': float(loss),
           'loss_smooth': float(info['loss_smooth']), 'loss_sparse_depth': float(info['loss_sparse_depth']),
           'loss_cos': float(info['loss_cos'])}
    if return_grads:
        res['grads'] = grads
        res['emb'] = emb.detach()
        res['ref'] = ref.detach()
    return res


# ----------------------------------------------------------------------------------------------------------------
# stage 2 of the source-domain preparation on the NLSPN back-end: the predictor heads (src/head_main.py:259-275, 437-480)
# ----------------------------------------------------------------------------------------------------------------
HEAD_TRAINED = tuple('%s.%s.%s' % (m, i, q) for m in ('proj', 'pred') for i in ('0', '1', '3') for q in ('weight', 'bias'))   # W:261-265
_HEAD_FLOAT = ('0.weight', '0.bias', '1.weight', '1.bias', '3.weight', '3.bias')


def _mlp_stage2(sd, name, x, train_bn, pr):
    """MLP head with its BatchNorm1d as stage 2 leaves it: running statistics present; train mode (batch statistics, running statistics
    and num_batches_tracked updated) for proj / pred, eval mode for proj_t (W:360-368 after convert_syncbn, head_main.py:275)"""
    h = pr.act(F.linear(x, pr.wgt(sd[name + '.0.weight']), sd[name + '.0.bias']))
    if train_bn:
        sd[name + '.1.num_batches_tracked'] += 1
    h = F.batch_norm(h, sd[name + '.1.running_mean'], sd[name + '.1.running_var'], sd[name + '.1.weight'], sd[name + '.1.bias'], train_bn, 0.1, 1e-5)
    return pr.act(F.linear(pr.act(F.relu(h)), pr.wgt(sd[name + '.3.weight']), sd[name + '.3.bias']))


def head_step(sd, state, image, sparse_depth, *, lr, pr=FP32, tau=0.999, weight_decay=0.0, return_grads=False):
    """One stage-2 step with forward 'head_meta_selfsup_seq_ema_reverse' = _rgbd_meta_contrast_prepare, mode [seq, reverse, ema]
    (M:1014-1060): both encoders under no_grad (BatchNorm2d in eval mode), proj_t <- EMA(proj) over its parameters (M:1314-1316),
    emb = pred(proj(fe6 of the zero image)), ref = proj_t(fe6 of the frame).detach(); loss 'prepare' = mean(2 - 2 cos) (E:524-540);
    Adam over proj.* and pred.* -- unlike MSG-CHN's variant only the INPUT of proj is detached (M:1057), so both heads train.
    `image` is the normalised network input (head_main.py:457-467 feeds the augmented image).  Mutates `sd`, `state`."""
    names = list(state.m.keys())
    with torch.no_grad():
        with running_batchnorm():
            fe6 = encoder(sd, image, sparse_depth, pr)[-1]
            fe6_z = encoder(sd, torch.zeros_like(image), sparse_depth, pr)[-1]
        for k in _HEAD_FLOAT:
            sd['proj_t.' + k].copy_(sd['proj_t.' + k] * tau + sd['proj.' + k] * (1.0 - tau))
    work = dict(sd)
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    emb = _mlp_stage2(work, 'pred', _mlp_stage2(work, 'proj', fe6_z.permute(0, 2, 3, 1).reshape(-1, 512), True, pr), True, pr)
    with torch.no_grad():
        ref = _mlp_stage2(sd, 'proj_t', fe6.permute(0, 2, 3, 1).reshape(-1, 512), False, pr)
    e = F.normalize(emb, dim=-1, p=2)
    r = F.normalize(ref, dim=-1, p=2)
    loss = (2 - 2 * (e * r).sum(-1)).mean()
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    O.adam_update(sd, grads, state, lr, weight_decay=weight_decay)
    res = {'loss': float(loss.detach())}
    if return_grads:
        res['grads'] = grads
        res['emb'] = emb.detach()
        res['ref'] = ref.detach()
    return res
