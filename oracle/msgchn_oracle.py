"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of ProxyTTA's per-frame
adaptation step for the MSG-CHN back-end.  It is the checker for the CUDA path, never the
thing shipped or measured: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.

Parity pinning: the reference ships no golden vectors and no runnable tests, so
this restatement is pinned against *outputs of the reference itself run in the build container*
(`oracle/gen_golden.py` imports /root/reference through `oracle/ref_shims.py`, feeds it the same
seeded synthetic checkpoint and frames, and commits the results under tests/golden/).
tests/test_oracle_golden.py checks this file against those fixtures (and, when /root/reference is
present, against the live reference).

Everything is functional over a flat state dict whose keys and shapes are exactly those of the
reference's `network_adapt.state_dict()` after `_prepare_head(mode)` (132 entries for
`meta_selfsup_seq_2layers_ema`), so the same `{'net': state_dict}` checkpoint feeds both.

Reference files restated here (paths relative to the reference root):
  N  = external_src/MSG_CHN/workspace/exp_msg_chn/network_exp_msg_chn_adapt.py
  W  = src/msg_chn_model_adapt.py
  E  = src/external_model_adapt.py
  L  = src/loss_utils.py
  U  = src/net_utils.py
  T  = src/tta_main.py
  V  = src/eval_utils.py
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# precision emulation hook (used only for tolerance studies: rounds a tensor to bf16 and back)
# ----------------------------------------------------------------------------------------------
class Precision:
    """`act(x)` is applied to every 32/128-channel activation and head matrix, `wgt(w)` to every conv / linear weight that the
    native path feeds to the tensor cores (the {1,2,3}->32 stems and the 32->1 prediction convs run in fp32 there, from the
    fp32 weights).  Identity in the fp32 oracle.  Because the rounding is applied with `.to(dtype)`, autograd rounds the gradient
    arriving at each of these tensors as well -- the stored gradient maps of the native backward."""

    def __init__(self, emulate=None, fuse_up2_min_pixels=6000):
        self.emulate = emulate
        self.dtype = {'bf16': torch.bfloat16, 'fp16': torch.float16}.get(emulate)
        # the native path adds the upsampled map of the cascade in the conv epilogue (ONE rounding of the sum) on maps its tcgen05 conv
        # takes (engine default tc_min_pixels), and in a separate pass (conv output rounded, then the sum rounded) on smaller maps
        self.fuse_up2_min_pixels = fuse_up2_min_pixels

    def conv_plus(self, y, add):
        """stored activation of `conv output y + upsampled map add` at the native rounding points"""
        if add is None:
            return self.act(y)
        if y.shape[0] * y.shape[2] * y.shape[3] < self.fuse_up2_min_pixels:
            y = self.act(y)
        return self.act(y + add)

    def act(self, x):
        if self.dtype is not None:
            return x.to(self.dtype).to(torch.float32)
        return x

    def wgt(self, w):
        if self.dtype is not None and not (w.dim() == 4 and min(w.shape[0], w.shape[1]) < 32):
            return w.to(self.dtype).to(torch.float32)
        return w


FP32 = Precision(None)


# ----------------------------------------------------------------------------------------------
# synthetic checkpoint / frames: input generators shared with bench.py, re-exported
# ----------------------------------------------------------------------------------------------
from tta_depth_completion_b200.synthetic import (make_synthetic_checkpoint, checkpoint_digest, synthetic_frame, DATASETS,  # noqa: E402,F401
                                                 load_fitted_checkpoint, get_checkpoint, fitted_checkpoint_available)


# ----------------------------------------------------------------------------------------------
# pre-processing (driver side): T:583-590, U:766-811
# ----------------------------------------------------------------------------------------------
def validity_map(sparse_depth):
    # T:583-586
    return torch.where(sparse_depth > 0, torch.ones_like(sparse_depth), sparse_depth)


def remove_outliers(sparse_depth, validity, kernel_size=7, threshold=1.5):
    # U:766-811
    max_value = 10 * torch.max(sparse_depth)
    filled = torch.where(validity <= 0, torch.full_like(sparse_depth, float(max_value)), sparse_depth)
    pad = kernel_size // 2
    filled = F.pad(filled, (pad, pad, pad, pad), mode='constant', value=float(max_value))
    min_values = -F.max_pool2d(-filled, kernel_size=kernel_size, stride=1, padding=0)
    clean = torch.where(min_values < sparse_depth - threshold,
                        torch.zeros_like(validity), torch.ones_like(validity))
    clean = validity * clean
    return sparse_depth * clean, clean


# ----------------------------------------------------------------------------------------------
# network (N:166-311, 463-557)
# ----------------------------------------------------------------------------------------------
def _conv(sd, name, x, pr, stride=1):
    return F.conv2d(x, pr.wgt(sd[name + '.weight']), sd.get(name + '.bias'), stride=stride, padding=1)


def _convT(sd, name, x, pr):
    return F.conv_transpose2d(x, pr.wgt(sd[name + '.weight']), sd[name + '.bias'], stride=2, padding=1,
                              output_padding=1)


def _up2(x):
    # Appendix B: align_corners=True
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)


def _enc_block(sd, prefix, x, pr, add=None):
    # ReLU, conv s2, ReLU, conv  (N:175-186); `add`: the upsampled map the caller adds to the result (N:201-209) -- the native path adds it
    # in the conv epilogue, BEFORE the one rounding of the stored activation, and the emulation rounds at the same point
    x = pr.act(_conv(sd, prefix + '.1', F.relu(x), pr, stride=2))
    return pr.conv_plus(_conv(sd, prefix + '.3', F.relu(x), pr), add)


def _init_block(sd, prefix, x, pr, add=None):
    # conv, ReLU, conv (N:171-173)
    x = pr.act(_conv(sd, prefix + '.0', x, pr))
    return pr.conv_plus(_conv(sd, prefix + '.2', F.relu(x), pr), add)


def rgb_encoder(sd, rgb, pr=FP32):
    # N:256-264
    x0 = _init_block(sd, 'rgb_encoder.init', rgb, pr)
    outs = [x0]
    for k in range(1, 5):
        outs.append(_enc_block(sd, 'rgb_encoder.enc%d' % k, outs[-1], pr))
    return outs


def depth_encoder(sd, prefix, inp, pre_x2=None, pre_x3=None, pre_x4=None, pr=FP32):
    # N:196-211
    x0 = _init_block(sd, prefix + '.init', inp, pr, None if pre_x4 is None else _up2(pre_x4))
    x1 = _enc_block(sd, prefix + '.enc1', x0, pr, None if pre_x3 is None else _up2(pre_x3))
    x2 = _enc_block(sd, prefix + '.enc2', x1, pr, None if pre_x2 is None else _up2(pre_x2))
    return x0, x1, x2


def _dec_block(sd, prefix, x, pr):
    # ReLU, convT s2, ReLU, conv (N:271-283)
    x = pr.act(_convT(sd, prefix + '.1', F.relu(x), pr))
    return pr.act(_conv(sd, prefix + '.3', F.relu(x), pr))


def depth_decoder(sd, prefix, pre_dx, pre_cx, pr=FP32):
    # N:297-311
    x2 = pr.act(pre_dx[2] + pre_cx[2])
    x1 = pr.act(pre_dx[1] + pre_cx[1])
    x0 = pr.act(pre_dx[0] + pre_cx[0])
    x3 = _dec_block(sd, prefix + '.dec2', x2, pr)
    x4 = _dec_block(sd, prefix + '.dec1', pr.act(x1 + x3), pr)
    h = pr.act(_conv(sd, prefix + '.prdct.1', F.relu(pr.act(x4 + x0)), pr))
    out = _conv(sd, prefix + '.prdct.3', F.relu(h), pr)
    return x2, x3, x4, out


def _bn_train(sd, name, x, training, momentum=0.1, eps=1e-5):
    """BatchNorm in the module's current mode; in train mode the running statistics in `sd` are
    updated in place, exactly as nn.BatchNorm does (this is what makes the reference update them
    twice per step, SURVEY.md gotcha 8)."""
    rm, rv = sd[name + '.running_mean'], sd[name + '.running_var']
    y = F.batch_norm(x, rm, rv, sd[name + '.weight'], sd[name + '.bias'], training, momentum, eps)
    if training:
        sd[name + '.num_batches_tracked'] += 1
    return y


def meta_layer(sd, x, training, pr=FP32):
    """conv1_rgb_meta: N:28-36 (Res_Conv, '2layers') or a plain Conv2d ('1layer', N:1066)."""
    if 'conv1_rgb_meta.weight' in sd:
        return pr.act(_conv(sd, 'conv1_rgb_meta', x, pr))
    p = 'conv1_rgb_meta.conv1_meta'
    h = pr.act(F.conv2d(x, pr.wgt(sd[p + '.0.0.weight']), None, padding=1))
    h = pr.act(F.leaky_relu(_bn_train(sd, p + '.0.1', h, training), 0.2))
    h = pr.act(_conv(sd, p + '.1', h, pr))
    h = _bn_train(sd, p + '.2', h, training)
    return pr.act(h + x)


def pyramid(d):
    # N:479,487,492  validity-normalised average pooling
    c = (d > 0).float()
    d14 = F.avg_pool2d(d, 4, 4) / (F.avg_pool2d(c, 4, 4) + 0.0001)
    d12 = F.avg_pool2d(d, 2, 2) / (F.avg_pool2d(c, 2, 2) + 0.0001)
    return d12, d14


def _cascade(sd, enc_c, d, d12, d14, with_decoder3, pr):
    # N:487-506 (real branch) / N:515-532 (zero branch, stops after depth_encoder3)
    enc14 = depth_encoder(sd, 'depth_encoder1', d14, pr=pr)
    dcd14 = depth_decoder(sd, 'depth_decoder1', enc14, enc_c[2:5], pr)
    p12 = _up2(dcd14[3])
    enc12 = depth_encoder(sd, 'depth_encoder2', torch.cat((d12, p12), 1), dcd14[0], dcd14[1], dcd14[2], pr)
    dcd12 = depth_decoder(sd, 'depth_decoder2', enc12, enc_c[1:4], pr)
    p11 = _up2(dcd12[3] + p12)
    enc11 = depth_encoder(sd, 'depth_encoder3', torch.cat((d, p11), 1), dcd12[0], dcd12[1], dcd12[2], pr)
    if not with_decoder3:
        return None, enc11
    dcd11 = depth_decoder(sd, 'depth_decoder3', enc11, enc_c[0:3], pr)
    return dcd11[3] + p11, enc11


def _mlp(sd, name, x, training, pr):
    # N:1089-1098
    h = pr.act(F.linear(x, pr.wgt(sd[name + '.0.weight']), sd[name + '.0.bias']))
    h = pr.act(F.relu(_bn_train(sd, name + '.1', h, training)))
    return pr.act(F.linear(h, pr.wgt(sd[name + '.3.weight']), sd[name + '.3.bias']))


def network_forward(sd, image, sparse_depth, training, pr=FP32):
    """network_adapt._rgbd_meta_contrast with mode = [adapt, reverse, seq, ema] (N:463-557), i.e.
    loss_type 'adapt_meta_selfsup_seq_ema_reverse'.  Train mode returns (output, emb, ref); eval mode
    returns output only."""
    d12, d14 = pyramid(sparse_depth)
    enc_c = rgb_encoder(sd, image, pr)
    enc_c[2] = meta_layer(sd, enc_c[2], training, pr)
    output, enc11 = _cascade(sd, enc_c, sparse_depth, d12, d14, True, pr)
    if not training:
        return output
    with torch.no_grad():
        enc_z = rgb_encoder(sd, torch.zeros_like(image), pr)
        enc_z[2] = meta_layer(sd, enc_z[2], training, pr)
        _, enc11_zero = _cascade(sd, enc_z, sparse_depth, d12, d14, False, pr)
    # N:551-554 ('ema', 'reverse', 'adapt'): rows are pixels of the /4 map in NHWC order
    z_zero = enc11_zero[2].permute(0, 2, 3, 1).reshape(-1, 32).detach()
    z_real = enc11[2].permute(0, 2, 3, 1).reshape(-1, 32)
    emb = _mlp(sd, 'pred', _mlp(sd, 'proj', z_zero, training, pr), training, pr)
    ref = _mlp(sd, 'proj', z_real, training, pr)
    return output, emb, ref


def model_forward(sd, image, sparse_depth, training, max_input_depth, pr=FP32):
    """ExternalModel_Adapt.forward (E:103-108: clamp) -> MsgChnModel_Adapt.forward (W:54-125).
    The pad-to-/16 + flip-pad ensembling (W:58-125) is restated for completeness."""
    if max_input_depth is not None:
        sparse_depth = torch.clamp(sparse_depth, 0, max_input_depth)
    h, w = image.shape[-2:]
    pt = (16 - h % 16) % 16
    prt = (16 - w % 16) % 16
    if pt == 0 and prt == 0:
        return network_forward(sd, image, sparse_depth, training, pr)
    image0 = F.pad(image, (0, prt, pt, 0, 0, 0)); sparse0 = F.pad(sparse_depth, (0, prt, pt, 0, 0, 0))
    image1 = F.pad(image, (prt, 0, 0, pt, 0, 0)); sparse1 = F.pad(sparse_depth, (prt, 0, 0, pt, 0, 0))
    out = network_forward(sd, torch.cat([image0, image1], 0), torch.cat([sparse0, sparse1], 0), training, pr)
    output = out[0] if training else out
    o0, o1 = torch.chunk(output, 2, 0)
    hh, ww = o0.shape[-2:]
    o0 = o0[:, :, pt:, :ww - prt]
    o1 = o1[:, :, :hh - pt, prt:]
    output = 0.5 * (o0 + o1)
    if training:
        return output, out[1], out[2]
    return output


# ----------------------------------------------------------------------------------------------
# losses: L:116-169, 624-638; E:371-441
# ----------------------------------------------------------------------------------------------
def smoothness_loss(predict, image):
    # L:139-169 with gradient_yx L:624-638
    pdx = predict[:, :, :, :-1] - predict[:, :, :, 1:]
    pdy = predict[:, :, :-1, :] - predict[:, :, 1:, :]
    idx = image[:, :, :, :-1] - image[:, :, :, 1:]
    idy = image[:, :, :-1, :] - image[:, :, 1:, :]
    wx = torch.exp(-torch.mean(torch.abs(idx), dim=1, keepdim=True))
    wy = torch.exp(-torch.mean(torch.abs(idy), dim=1, keepdim=True))
    return torch.mean(wx * torch.abs(pdx)) + torch.mean(wy * torch.abs(pdy))


def sparse_depth_loss(src, tgt, w):
    # L:116-137 (no epsilon: an empty frame gives NaN in the reference too)
    delta = torch.abs(tgt - src)
    loss = torch.sum(w * delta, dim=[1, 2, 3])
    return torch.mean(loss / torch.sum(w, dim=[1, 2, 3]))


def adapt_loss(input_rgb, output_depth, sparse_depth, validity, embedding, reference,
               w_sd, w_sm, w_cos, max_input_depth=None):
    # E:191-203 (clamp) + E:371-441
    if max_input_depth is not None:
        sparse_depth = torch.clamp(sparse_depth, 0, max_input_depth)
    l_sm = smoothness_loss(output_depth, input_rgb)
    l_sd = sparse_depth_loss(output_depth, sparse_depth, validity)
    e = F.normalize(embedding, dim=-1, p=2)
    r = F.normalize(reference, dim=-1, p=2)
    l_cos = (2 - 2 * (e * r).sum(dim=-1)).mean()
    if l_cos < 0.3:           # E:424 (a host sync in the reference)
        w_cos = 0
    loss = w_sd * l_sd + w_sm * l_sm + w_cos * l_cos
    return loss, {'loss': loss, 'loss_smooth': l_sm, 'loss_sparse_depth': l_sd, 'loss_cos': l_cos}


# ----------------------------------------------------------------------------------------------
# adapted parameters + Adam: W:392-396; T:341-346,633 (torch.optim.Adam, amsgrad False, wd 0)
# ----------------------------------------------------------------------------------------------
_FLOAT_STATE = ('weight', 'bias')


def adapt_parameter_names(sd, mode='meta'):
    # W:392-396: every *parameter* (not buffer) whose name contains 'meta'
    if mode != 'meta':
        raise NotImplementedError(mode)
    return [k for k in sd if 'meta' in k and k.rsplit('.', 1)[-1] in _FLOAT_STATE]


class AdamState:
    def __init__(self, names, sd):
        self.step = 0
        self.m = {k: torch.zeros_like(sd[k]) for k in names}
        self.v = {k: torch.zeros_like(sd[k]) for k in names}


def adam_update(sd, grads, state, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam single-tensor formulation."""
    state.step += 1
    b1, b2 = betas
    bc1 = 1 - b1 ** state.step
    bc2 = 1 - b2 ** state.step
    for k, g in grads.items():
        p = sd[k]
        if weight_decay != 0:
            g = g + weight_decay * p
        state.m[k].mul_(b1).add_(g, alpha=1 - b1)
        state.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (state.v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        p.data.addcdiv_(state.m[k], denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------------------------
# one full TTA step: T:583-633
# ----------------------------------------------------------------------------------------------
def tta_step(sd, state, image, sparse_depth, *, lr, w_sd=1.0, w_sm=1.0, w_cos=0.1,
             max_input_depth=80.0, normalize=lambda im: im / 255.0, pr=FP32, return_grads=False):
    """`image` is the raw [0,255] image; the network sees `normalize(image)` (T:595-604, 610), the
    smoothness loss sees the raw one (T:620; SURVEY.md gotcha 7).  Mutates `sd` (adapted tensors,
    BN buffers) and `state`.  Returns a dict of everything the parity tests compare."""
    names = list(state.m.keys())
    v = validity_map(sparse_depth)
    d_f, v_f = remove_outliers(sparse_depth, v)
    work = dict(sd)
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    out, emb, ref = model_forward(work, normalize(image), d_f, True, max_input_depth, pr)
    loss, info = adapt_loss(image, out, d_f, v_f, emb, ref, w_sd, w_sm, w_cos, max_input_depth)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    adam_update(sd, grads, state, lr)
    res = {
        'validity': v_f, 'sparse_depth': d_f, 'output_depth': out.detach(),
        'loss': float(loss), 'loss_smooth': float(info['loss_smooth']),
        'loss_sparse_depth': float(info['loss_sparse_depth']), 'loss_cos': float(info['loss_cos']),
    }
    if return_grads:
        res['grads'] = grads
        res['emb'] = emb.detach()
        res['ref'] = ref.detach()
    return res


# ----------------------------------------------------------------------------------------------
# source-domain preparation: stage 1 src/init_main.py:448-572, stage 2 src/head_main.py:415-541
# ----------------------------------------------------------------------------------------------
def l2_loss(src, tgt, w):
    # L:266-287
    loss = (src - tgt) ** 2
    loss = torch.sum(w * loss, dim=[1, 2, 3]) / torch.sum(w, dim=[1, 2, 3])
    return torch.mean(loss)


def supervised_loss(output_depth, ground_truth, max_predict_depth=100.0):
    # W:224-264 (w_scale0 = 1, the two coarser scales carry weight 0)
    gt = torch.clamp(ground_truth, min=0.0, max=max_predict_depth)
    v = torch.where(gt > 0, torch.ones_like(gt), gt)
    return l2_loss(output_depth, gt, v)


def init_forward(sd, image, sparse_depth, max_input_depth, pr=FP32):
    """network_adapt._rgbd_meta_contrast_init (N:559-607) behind ExternalModel_Adapt.forward's clamp (E:103-108): the real branch only,
    meta-layer BatchNorm in train mode; returns output_d11 (the only scale the loss weights, W:240-242)."""
    if max_input_depth is not None:
        sparse_depth = torch.clamp(sparse_depth, 0, max_input_depth)
    d12, d14 = pyramid(sparse_depth)
    enc_c = rgb_encoder(sd, image, pr)
    enc_c[2] = meta_layer(sd, enc_c[2], True, pr)
    output, _ = _cascade(sd, enc_c, sparse_depth, d12, d14, True, pr)
    return output


def init_step(sd, state, image, sparse_depth, ground_truth, *, lr, max_input_depth=80.0, max_predict_depth=100.0,
              normalize=lambda im: im / 255.0, pr=FP32, weight_decay=0.0, return_grads=False):
    """One stage-1 step (src/init_main.py:482-522): forward 'init_meta_seq_ema', masked L2 against the ground truth, backward to the
    meta tensors, Adam.  Mutates `sd` and `state`."""
    names = list(state.m.keys())
    work = dict(sd)
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    out = init_forward(work, normalize(image), sparse_depth, max_input_depth, pr)
    loss = supervised_loss(out, ground_truth, max_predict_depth)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    adam_update(sd, grads, state, lr, weight_decay=weight_decay)
    res = {'loss': float(loss.detach()), 'output_depth': out.detach()}
    if return_grads:
        res['grads'] = grads
    return res


HEAD_TRAINED = ('pred.0.weight', 'pred.0.bias', 'pred.1.weight', 'pred.1.bias', 'pred.3.weight', 'pred.3.bias')


def head_step(sd, state, image, sparse_depth, *, lr, max_input_depth=80.0, normalize=lambda im: im / 255.0, pr=FP32, tau=0.999,
              weight_decay=0.0, return_grads=False):
    """One stage-2 step (src/head_main.py:437-480) with forward 'head_meta_selfsup_seq_ema_reverse' = _rgbd_meta_contrast_head with
    mode [reverse, seq, ema] (N:609-699): the whole network under no_grad on the frame and on the zero image (meta-layer BatchNorm in
    train mode: running statistics updated twice), proj_t <- EMA(proj) (N:701-703), emb = pred(proj(z_zero).detach()),
    ref = proj(z_real).detach(); loss 'prepare' = mean(2 - 2 cos) (E:524-540).  head_main.py:268 gives proj.* and pred.* to Adam, but
    only pred receives gradients (torch.optim.Adam skips tensors without one).  `state` covers HEAD_TRAINED.  Mutates `sd`, `state`."""
    names = list(state.m.keys())
    image = normalize(image)
    if max_input_depth is not None:
        sparse_depth = torch.clamp(sparse_depth, 0, max_input_depth)
    with torch.no_grad():
        d12, d14 = pyramid(sparse_depth)
        enc_c = rgb_encoder(sd, image, pr)
        enc_c[2] = meta_layer(sd, enc_c[2], True, pr)
        _, enc11 = _cascade(sd, enc_c, sparse_depth, d12, d14, False, pr)
        enc_z = rgb_encoder(sd, torch.zeros_like(image), pr)
        enc_z[2] = meta_layer(sd, enc_z[2], True, pr)
        _, enc11_zero = _cascade(sd, enc_z, sparse_depth, d12, d14, False, pr)
        for k in [k for k in sd if k.startswith('proj_t.') and k.rsplit('.', 1)[-1] in _FLOAT_STATE]:
            sd[k].copy_(sd[k] * tau + sd['proj.' + k[7:]] * (1.0 - tau))
        z_zero = enc11_zero[2].permute(0, 2, 3, 1).reshape(-1, 32)
        z_real = enc11[2].permute(0, 2, 3, 1).reshape(-1, 32)
        pz = _mlp(sd, 'proj', z_zero, True, pr)
    work = dict(sd)
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    emb = _mlp(work, 'pred', pz.detach(), True, pr)
    with torch.no_grad():
        ref = _mlp(sd, 'proj', z_real, True, pr)
    e = F.normalize(emb, dim=-1, p=2)
    r = F.normalize(ref, dim=-1, p=2)
    loss = (2 - 2 * (e * r).sum(-1)).mean()
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    adam_update(sd, grads, state, lr, weight_decay=weight_decay)
    res = {'loss': float(loss.detach())}
    if return_grads:
        res['grads'] = grads
        res['emb'] = emb.detach()
        res['ref'] = ref.detach()
    return res


# ----------------------------------------------------------------------------------------------
# evaluation metrics: V:117-175 (torch variants used by T:760-798), after the depth-range mask
# ----------------------------------------------------------------------------------------------
def eval_metrics(output_depth, ground_truth, min_depth, max_depth):
    """T:773-798 with V:117-175: mask = gt>0 and min<=gt<=max; MAE / RMSE on 1000*depth (mm),
    iMAE / iRMSE on 0.001*depth (1/km), eps 1e-9 (V:24)."""
    eps = 1e-9
    mask = (ground_truth > 0) & ~(ground_truth < min_depth) & ~(ground_truth > max_depth)
    o = output_depth[mask]
    g = ground_truth[mask]
    mae = torch.mean(torch.abs(1000.0 * g - 1000.0 * o))
    rmse = torch.sqrt(torch.mean((1000.0 * g - 1000.0 * o) ** 2))
    imae = torch.mean(torch.abs(1.0 / (0.001 * g + eps) - 1.0 / (0.001 * o + eps)))
    irmse = torch.sqrt(torch.mean((1.0 / (0.001 * g + eps) - 1.0 / (0.001 * o + eps)) ** 2))
    return {'mae': float(mae), 'rmse': float(rmse), 'imae': float(imae), 'irmse': float(irmse)}
