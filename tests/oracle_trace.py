"""Test helper: the oracle's training forward + loss + backward with named intermediates, using the same names
as the native engine's tensor registry, so that a parity failure can be localised to a block."""
import torch
import torch.nn.functional as F

from oracle import msgchn_oracle as O


def trace_step(sd, image_raw, sparse_raw, cap, w_sd=1.0, w_sm=1.0, w_cos=0.1, pr=O.FP32):
    """Returns (T, G, losses, grads): forward tensors, gradients wrt selected intermediates, loss scalars and the
    gradients of the adapted tensors.  Does NOT update sd's adapted tensors (BN buffers are updated, as in a real step)."""
    sd = dict(sd)
    names = O.adapt_parameter_names(sd, 'meta')
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        sd[k] = leaves[k]
    T = {}

    def keep(name, t):
        if t.requires_grad:
            t.retain_grad()
        T[name] = t
        return t

    v = O.validity_map(sparse_raw)
    d_f, v_f = O.remove_outliers(sparse_raw, v)
    T['filtered_depth'], T['filtered_validity'] = d_f, v_f
    image = image_raw / 255.0
    d = torch.clamp(d_f, 0, cap)
    d12, d14 = O.pyramid(d)
    T['depth_clamped'], T['d12'], T['d14'] = d, d12, d14

    def cascade(tag, enc_c, with_dec3):
        e1 = O.depth_encoder(sd, 'depth_encoder1', d14, pr=pr)
        for i, t in enumerate(e1):
            keep('%s.e1.x%d' % (tag, i), t)
        dc1 = O.depth_decoder(sd, 'depth_decoder1', e1, enc_c[2:5], pr)
        for nm, t in zip(('x2', 'x3', 'x4', 'out'), dc1):
            keep('%s.d1.%s' % (tag, nm), t)
        p12 = keep(tag + '.p12', O._up2(dc1[3]))
        e2 = O.depth_encoder(sd, 'depth_encoder2', torch.cat((d12, p12), 1), dc1[0], dc1[1], dc1[2], pr)
        for i, t in enumerate(e2):
            keep('%s.e2.x%d' % (tag, i), t)
        # what the native path stores of x0 / x1 (their raw values are read by nothing): the ReLU copies and the decoder's sums
        for i in (0, 1):
            T['%s.e2.x%dr' % (tag, i)] = F.relu(e2[i])
            T['%s.d2.x%d' % (tag, i)] = pr.act(e2[i] + enc_c[1 + i])
        dc2 = O.depth_decoder(sd, 'depth_decoder2', e2, enc_c[1:4], pr)
        for nm, t in zip(('x2', 'x3', 'x4', 'out'), dc2):
            keep('%s.d2.%s' % (tag, nm), t)
        p11 = keep(tag + '.p11', O._up2(dc2[3] + p12))
        e3 = O.depth_encoder(sd, 'depth_encoder3', torch.cat((d, p11), 1), dc2[0], dc2[1], dc2[2], pr)
        for i, t in enumerate(e3):
            keep('%s.e3.x%d' % (tag, i), t)
        for i in (0, 1):
            T['%s.e3.x%dr' % (tag, i)] = F.relu(e3[i])
            if with_dec3:
                T['%s.d3.x%d' % (tag, i)] = pr.act(e3[i] + enc_c[i])
        if not with_dec3:
            return None, e3
        dc3 = O.depth_decoder(sd, 'depth_decoder3', e3, enc_c[0:3], pr)
        for nm, t in zip(('x2', 'x3', 'x4', 'out'), dc3):
            keep('%s.d3.%s' % (tag, nm), t)
        return keep(tag + '.output', dc3[3] + p11), e3

    enc_c = O.rgb_encoder(sd, image, pr)
    for i, t in enumerate(enc_c):
        T['real.c%d' % i if i != 2 else 'real.c2raw'] = t
    enc_c[2] = keep('real.c2', O.meta_layer(sd, enc_c[2], True, pr))
    output, e3 = cascade('real', enc_c, True)
    with torch.no_grad():
        enc_z = O.rgb_encoder(sd, torch.zeros_like(image), pr)
        for i, t in enumerate(enc_z):
            T['zc%d' % i] = t
        enc_z[2] = O.meta_layer(sd, enc_z[2], True, pr)
        T['zero.c2'] = enc_z[2]
        _, e3z = cascade('zero', enc_z, False)
    z_zero = e3z[2].permute(0, 2, 3, 1).reshape(-1, 32).detach()
    z_real = e3[2].permute(0, 2, 3, 1).reshape(-1, 32)
    emb = O._mlp(sd, 'pred', O._mlp(sd, 'proj', z_zero, True, pr), True, pr)
    ref = keep('ref', O._mlp(sd, 'proj', z_real, True, pr))
    T['emb'] = emb
    loss, info = O.adapt_loss(image_raw, output, d_f, v_f, emb, ref, w_sd, w_sm, w_cos, cap)
    loss.backward()
    G = {}
    for name, key in (('g_output', 'real.output'), ('g_p11', 'real.p11'), ('g_p12', 'real.p12'), ('g_out14', 'real.d1.out'),
                      ('g_c2', 'real.c2'), ('g_ref', 'ref'), ('g_e3x2', 'real.e3.x2'), ('g_out12', 'real.d2.out')):
        g = T[key].grad
        G[name] = g.detach() if g is not None else None
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])) for k in names}
    losses = {'loss': float(loss), 'loss_smooth': float(info['loss_smooth']),
              'loss_sparse_depth': float(info['loss_sparse_depth']), 'loss_cos': float(info['loss_cos'])}
    T = {k: t.detach() for k, t in T.items()}
    return T, G, losses, grads


def to_nchw(t):
    """engine tensor (NHWC bf16 / [N,H,W,1] fp32) -> NCHW float CPU"""
    t = t.float().cpu()
    if t.dim() == 4:
        return t.permute(0, 3, 1, 2).contiguous()
    return t
