"""Seeded synthetic workload of SURVEY.md section 8d: stand-in checkpoints (the reference ships none: README.md:144-149 points at
Google Drive) and frames.  Input generators only -- no arithmetic of the TTA step lives here.  Used by bench.py, __graft_entry__.smoke(),
the tests and (re-exported) by the CPU oracles, so that every side is fed the same bits.

Key set / shapes of the MSG-CHN state dict follow external_src/MSG_CHN/workspace/exp_msg_chn/network_exp_msg_chn_adapt.py
(N:166-335 network, N:1022-1087 `_prepare_head`)."""
import math
import os
from collections import OrderedDict

import torch


# ----------------------------------------------------------------------------------------------
# synthetic checkpoint / frames
# ----------------------------------------------------------------------------------------------
def _conv_entry(sd, g, name, cout, cin, bias=True, transposed=False):
    # N:188-194  xavier_normal_ weights, bias 0.01 (encoders / decoders)
    shape = (cin, cout, 3, 3) if transposed else (cout, cin, 3, 3)
    fan_in, fan_out = cin * 9, cout * 9
    if transposed:
        fan_in, fan_out = cout * 9, cin * 9
    std = math.sqrt(2.0 / (fan_in + fan_out))
    sd[name + '.weight'] = torch.randn(shape, generator=g) * std
    if bias:
        sd[name + '.bias'] = torch.full((cout,), 0.01)


def _bn_entry(sd, g, name, c):
    # affine / running stats perturbed away from the (1, 0, 0, 1) init so that tests see them
    sd[name + '.weight'] = 1.0 + 0.1 * torch.randn(c, generator=g)
    sd[name + '.bias'] = 0.05 * torch.randn(c, generator=g)
    sd[name + '.running_mean'] = 0.05 * torch.randn(c, generator=g)
    sd[name + '.running_var'] = 1.0 + 0.1 * torch.rand(c, generator=g)
    sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def _linear_entry(sd, g, name, cout, cin):
    bound = 1.0 / math.sqrt(cin)       # nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    sd[name + '.weight'] = (torch.rand((cout, cin), generator=g) * 2 - 1) * bound
    sd[name + '.bias'] = (torch.rand((cout,), generator=g) * 2 - 1) * bound


def _mlp_entries(sd, g, name, dim, out, hidden):
    # N:1089-1098  Linear -> BatchNorm1d -> ReLU -> Linear
    _linear_entry(sd, g, name + '.0', hidden, dim)
    _bn_entry(sd, g, name + '.1', hidden)
    _linear_entry(sd, g, name + '.3', out, hidden)


def make_synthetic_checkpoint(seed=0, prepare_mode='meta_selfsup_seq_2layers_ema'):
    """Seeded stand-in for the Google-Drive checkpoints (none is in the reference tree).  Key set
    and shapes follow N:166-335 (network) and N:1022-1087 (`_prepare_head`)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def encoder(prefix, cin, n_enc):
        _conv_entry(sd, g, prefix + '.init.0', 32, cin)
        _conv_entry(sd, g, prefix + '.init.2', 32, 32)
        for k in range(1, n_enc + 1):
            _conv_entry(sd, g, '%s.enc%d.1' % (prefix, k), 32, 32)
            _conv_entry(sd, g, '%s.enc%d.3' % (prefix, k), 32, 32)

    def decoder(prefix):
        for blk in ('dec2', 'dec1'):
            _conv_entry(sd, g, '%s.%s.1' % (prefix, blk), 32, 32, transposed=True)
            _conv_entry(sd, g, '%s.%s.3' % (prefix, blk), 32, 32)
        _conv_entry(sd, g, prefix + '.prdct.1', 32, 32)
        _conv_entry(sd, g, prefix + '.prdct.3', 1, 32)

    encoder('rgb_encoder', 3, 4)
    encoder('depth_encoder1', 1, 2)
    decoder('depth_decoder1')
    encoder('depth_encoder2', 2, 2)
    decoder('depth_decoder2')
    encoder('depth_encoder3', 2, 2)
    decoder('depth_decoder3')
    if 'selfsup' in prepare_mode:
        _mlp_entries(sd, g, 'proj', 32, 512, 512)
        if 'ema' in prepare_mode:
            for k in [k for k in sd if k.startswith('proj.')]:
                sd['proj_t.' + k[5:]] = sd[k].clone()
        _mlp_entries(sd, g, 'pred', 512, 512, 512)
    if 'meta' in prepare_mode and 'seq' in prepare_mode:
        if '1layer' in prepare_mode:
            # N:1066-1068  Conv2d(32,32,3,1,1), kaiming_normal fan_out
            sd['conv1_rgb_meta.weight'] = torch.randn((32, 32, 3, 3), generator=g) * math.sqrt(2.0 / (32 * 9))
            sd['conv1_rgb_meta.bias'] = (torch.rand((32,), generator=g) * 2 - 1) / math.sqrt(32 * 9)
        elif '2layers' in prepare_mode:
            # N:28-36, 1073  Res_Conv(32,128,3,1,1)
            p = 'conv1_rgb_meta.conv1_meta'
            bound = 1.0 / math.sqrt(32 * 9)
            sd[p + '.0.0.weight'] = (torch.rand((128, 32, 3, 3), generator=g) * 2 - 1) * bound
            _bn_entry(sd, g, p + '.0.1', 128)
            bound = 1.0 / math.sqrt(128 * 9)
            sd[p + '.1.weight'] = (torch.rand((32, 128, 3, 3), generator=g) * 2 - 1) * bound
            sd[p + '.1.bias'] = (torch.rand((32,), generator=g) * 2 - 1) * bound
            _bn_entry(sd, g, p + '.2', 32)
        else:
            raise NotImplementedError(prepare_mode)
    return sd


def checkpoint_digest(sd):
    """Order-independent fingerprint of a state dict (guards the 'same seed -> same checkpoint on
    the GPU box' assumption the fixtures rely on)."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + 0.5 * float(v.abs().sum()) + 1e-3 * v.numel()
    return tot


DATASETS = {
    # name: (sampling density, depth cap, dense depth surface)  -- SURVEY.md §8d
    'kitti': (0.05, 80.0),
    'void': (0.005, 8.0),
}


def synthetic_frame(seq_seed, t, n, h, w, dataset='kitti', outlier_fraction=0.01, depth_scale=1.0):
    """Frame t of synthetic sequence `seq_seed`: image in [0,255]; sparse depth = smooth
    surface x Bernoulli(p), with ~1 % of the samples pushed +5 m so the outlier filter has work.
    depth_scale != 1 models a test-domain shift (the sensor reads depth_scale x the surface the source-domain network was
    fitted on): the sparse samples are scaled, the returned dense ground truth is scaled with them."""
    p, cap = DATASETS[dataset]
    g = torch.Generator().manual_seed(1000 * seq_seed + t)
    yy = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1)
    xx = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w)
    # smooth texture + +-4 grey levels of noise: i.i.d. uniform [0,255] pixels would make the
    # edge-aware weights exp(-|dI|) underflow to 0 and the smoothness loss vanish
    ph = torch.tensor([0.0, 1.3, 2.1]).view(1, 3, 1, 1)
    image = 127.0 + 100.0 * torch.sin(xx / 31.0 + ph + 0.05 * t) * torch.cos(yy / 17.0 + 0.5 * ph)
    image = image + 8.0 * (torch.rand((n, 3, h, w), generator=g) - 0.5)
    image = image.clamp(0.0, 255.0).contiguous()
    if dataset == 'kitti':
        dense = 5.0 + 70.0 * (1.0 - yy / h) + 2.0 * torch.sin((xx + 3.0 * t) / 97.0)
    else:
        dense = 0.5 + 4.0 * (yy / h) + 0.3 * torch.sin((xx + 3.0 * t) / 53.0)
    dense = (dense * depth_scale if depth_scale != 1.0 else dense).expand(n, 1, h, w).contiguous()
    mask = (torch.rand((n, 1, h, w), generator=g) < p).float()
    out = (torch.rand((n, 1, h, w), generator=g) < outlier_fraction).float()
    sparse = (dense + 5.0 * out * (cap / 80.0)) * mask
    return image, sparse, dense



# ----------------------------------------------------------------------------------------------
# fitted checkpoints (tests/golden/ckpt_*.pt, written by oracle/make_fitted_checkpoint.py with the real reference)
# ----------------------------------------------------------------------------------------------
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def fitted_checkpoint_available(name):
    return os.path.exists(os.path.join(GOLDEN_DIR, 'ckpt_%s.pt' % name))


def load_fitted_checkpoint(name):
    """name: 'kitti_2layers_a' | 'kitti_2layers_b' | 'void_1layer_a' | 'void_1layer_b' -> state dict (fp32 CPU tensors).
    `_a`: heads fitted for a few steps (loss_cos stays above the 0.3 gate), `_b`: heads fully fitted (the gate fires)."""
    blob = torch.load(os.path.join(GOLDEN_DIR, 'ckpt_%s.pt' % name), map_location='cpu', weights_only=False)
    if 'net' in blob:
        return OrderedDict((k, v.clone()) for k, v in blob['net'].items())
    base = torch.load(os.path.join(GOLDEN_DIR, blob['base']), map_location='cpu', weights_only=False)['net']
    sd = OrderedDict((k, v.clone()) for k, v in base.items())
    for k, v in blob['net_delta'].items():
        sd[k] = v.clone()
    return sd


def get_checkpoint(spec, prepare_mode):
    """spec: int seed -> randomly initialised stand-in; str -> fitted checkpoint of that name"""
    if isinstance(spec, str):
        return load_fitted_checkpoint(spec)
    return make_synthetic_checkpoint(spec, prepare_mode)


# ================================================================================================================
# NLSPN back-end (key order of the reference module: external_src/NLSPN/src/model/nlspnmodel_adapt.py M:384-448, heads M:1338-1374)
# ================================================================================================================
NLSPN_PREPARE_MODE = 'meta_selfsup_seq_1layer_ema'
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
RESNET34_LAYERS = (('conv2', 64, 64, 3, 1), ('conv3', 64, 128, 4, 2), ('conv4', 128, 256, 6, 2), ('conv5', 256, 512, 3, 2))


# ----------------------------------------------------------------------------------------------------------------
# synthetic checkpoint: seeded, so the GPU box regenerates the identical state dict
# ----------------------------------------------------------------------------------------------------------------
def _nl_conv(sd, g, name, cout, cin, k=3, bias=False, transposed=False, gain=1.0):
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    std = gain * math.sqrt(2.0 / (cin * k * k))
    sd[name + '.weight'] = torch.randn(shape, generator=g) * std
    if bias:
        sd[name + '.bias'] = 0.02 * torch.randn(cout, generator=g)


def _nl_bn(sd, g, name, c):
    sd[name + '.weight'] = 1.0 + 0.1 * torch.randn(c, generator=g)
    sd[name + '.bias'] = 0.05 * torch.randn(c, generator=g)
    sd[name + '.running_mean'] = 0.05 * torch.randn(c, generator=g)
    sd[name + '.running_var'] = 1.0 + 0.1 * torch.rand(c, generator=g)
    sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def make_nlspn_checkpoint(seed=0, prepare_mode=NLSPN_PREPARE_MODE):
    """Key order follows the reference module's registration order (M:384-448 then `_prepare_head` M:1338-1374)."""
    if 'meta' not in prepare_mode or '1layer' not in prepare_mode or 'ema' not in prepare_mode:
        raise NotImplementedError(prepare_mode)
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    _nl_conv(sd, g, 'conv1_rgb.0', 48, 3, bias=True)
    _nl_conv(sd, g, 'conv1_dep.0', 16, 1, bias=True, gain=0.05)      # depth in metres (up to 80) -> O(1) features
    for name, cin, cout, blocks, stride in RESNET34_LAYERS:
        for b in range(blocks):
            p = '%s.%d' % (name, b)
            _nl_conv(sd, g, p + '.conv1', cout, cin if b == 0 else cout)
            _nl_bn(sd, g, p + '.bn1', cout)
            _nl_conv(sd, g, p + '.conv2', cout, cout, gain=0.5)
            _nl_bn(sd, g, p + '.bn2', cout)
            if b == 0 and stride != 1:
                _nl_conv(sd, g, p + '.downsample.0', cout, cin, k=1)
                _nl_bn(sd, g, p + '.downsample.1', cout)
    _nl_conv(sd, g, 'conv6.0', 512, 512)
    _nl_bn(sd, g, 'conv6.1', 512)
    for name, cin, cout in (('dec5', 512, 256), ('dec4', 768, 128), ('dec3', 384, 64), ('dec2', 192, 64)):
        _nl_conv(sd, g, name + '.0', cout, cin, transposed=True)
        _nl_bn(sd, g, name + '.1', cout)
    _nl_conv(sd, g, 'id_dec1.0', 64, 128)
    _nl_bn(sd, g, 'id_dec1.1', 64)
    _nl_conv(sd, g, 'id_dec0.0', 1, 128, bias=True)
    sd['id_dec0.0.bias'] += 8.0                    # initial depth of a few metres, so the propagated depth is not clamped to 0
    _nl_conv(sd, g, 'gd_dec1.0', 64, 128)
    _nl_bn(sd, g, 'gd_dec1.1', 64)
    _nl_conv(sd, g, 'gd_dec0.0', 8, 128, bias=True)
    _nl_conv(sd, g, 'cf_dec1.0', 32, 128)
    _nl_bn(sd, g, 'cf_dec1.1', 32)
    _nl_conv(sd, g, 'cf_dec0.0', 1, 96, bias=True)
    # prop_layer (M:219-247): conv_offset_aff is zero-initialised by the reference, which makes the propagation the identity
    # -> seeded values (offsets ~1 px, affinities small), SURVEY.md 8c
    scale = torch.cat((torch.full((16,), 0.05), torch.full((8,), 0.004))).view(24, 1, 1, 1)
    sd['prop_layer.aff_scale_const'] = torch.full((1,), 0.5 * 8)
    sd['prop_layer.w'] = torch.ones((1, 1, 3, 3))
    sd['prop_layer.b'] = torch.zeros(1)
    sd['prop_layer.w_conf'] = torch.ones((1, 1, 1, 1))
    sd['prop_layer.conv_offset_aff.weight'] = torch.randn((24, 8, 3, 3), generator=g) * scale
    sd['prop_layer.conv_offset_aff.bias'] = torch.randn((24,), generator=g) * 0.05
    # heads (M:1338-1343): proj, proj_t = deepcopy(proj), pred
    _mlp_entries(sd, g, 'proj', 512, 1024, 1024)
    for k in [k for k in sd if k.startswith('proj.')]:
        sd['proj_t.' + k[5:]] = sd[k].clone()
    _mlp_entries(sd, g, 'pred', 1024, 1024, 1024)
    # meta layer (M:1370-1374): Conv2d(48,48,3,1,1)
    _nl_conv(sd, g, 'conv1_rgb_meta', 48, 48, bias=True, gain=0.7)
    return sd


def normalize_image_imagenet(image):
    """bash/adapt/adapt_nlspn_vkitti.sh:25-28: ImageNet statistics on the [0,1] image (T:595-604)"""
    mean = torch.tensor(IMAGENET_MEAN, dtype=image.dtype).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=image.dtype).view(1, 3, 1, 1)
    return (image / 255.0 - mean) / std


