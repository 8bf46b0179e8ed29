"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/nlspn_prop_*.pt by running the REAL reference NLSPN propagation
module (/root/reference/external_src/NLSPN/src/model/nlspnmodel_adapt.py:189-373, class NLSPN) on seeded inputs.

    python oracle/gen_golden_nlspn.py      # needs /root/reference; CPU, a few seconds

The reference imports `ModulatedDeformConvFunction` from its DCN CUDA extension, which has no CPU path
(modulated_deform_conv.h:45 "Not implemented on the CPU"); here that one import is resolved to
oracle.nlspn_prop_oracle.MDConvFn (a restatement of the extension's kernels that agrees with torchvision's deform_conv2d
to 1e-15 in fp64 and, on the GPU box, with the extension itself built by oracle/build_ref_dcn.py).  Everything else --
offset layout, TGASS normalisation, confidence gather with the legacy grid offset, reference affinity, input-preserving
18-step loop -- is the reference's own Python.  `conv_offset_aff` is zero-initialised by the reference (:223-224), which makes
propagation the identity; it is filled with seeded Gaussian values here.

Each fixture stores the inputs, the conv_offset_aff parameters, the intermediate (offset, affinity), the propagated
features of iterations 1, 9 and 18, and the gradients of a seeded linear functional of the output with respect to the
initial depth, the guidance, the confidence, the conv output and the conv parameters."""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import nlspn_prop_oracle as P       # noqa: E402

REF_MODEL_DIR = '/root/reference/external_src/NLSPN/src/model'
GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

CASES = [
    dict(name='nlspn_prop_1x24x40', seed=11, n=1, h=24, w=40, prop_time=18),
    dict(name='nlspn_prop_2x17x23', seed=12, n=2, h=17, w=23, prop_time=18),
]


def load_reference_nlspn():
    shim = types.ModuleType('modulated_deform_conv_func')
    shim.ModulatedDeformConvFunction = P.MDConvFn
    sys.modules['modulated_deform_conv_func'] = shim
    sys.path.insert(0, REF_MODEL_DIR)
    import nlspnmodel_adapt                     # the reference's own module
    return nlspnmodel_adapt.NLSPN


def run_case(NLSPN, case):
    args = types.SimpleNamespace(prop_time=case['prop_time'], affinity='TGASS', affinity_gamma=0.5, conf_prop=True, legacy=True,
                                 preserve_input=True)
    g = torch.Generator().manual_seed(case['seed'])
    n, h, w = case['n'], case['h'], case['w']
    mod = NLSPN(args, 8, 1, 3, 3)
    with torch.no_grad():
        # offsets (channels 0-15) ~1.7 px so that many samples leave the image; affinities (16-23) small enough that the
        # 18-step iteration stays bounded, as it is for trained networks
        scale = torch.cat((torch.full((16,), 0.05), torch.full((8,), 0.01))).view(24, 1, 1, 1)
        mod.conv_offset_aff.weight.copy_(torch.randn(mod.conv_offset_aff.weight.shape, generator=g) * scale)
        mod.conv_offset_aff.bias.copy_(torch.randn(mod.conv_offset_aff.bias.shape, generator=g) * 0.05)
    feat_init, sparse, _, confidence = P.synthetic_prop_inputs(case['seed'], n, h, w)
    guidance = torch.randn((n, 8, h, w), generator=g) * 4.0
    feat_init.requires_grad_(True); guidance.requires_grad_(True); confidence.requires_grad_(True)
    kept = {}

    def hook(m, i, o):
        o.retain_grad()
        kept['offset_aff'] = o
    mod.conv_offset_aff.register_forward_hook(hook)
    y, feats, offset, aff, scale = mod(feat_init, guidance, confidence, sparse)
    probe = torch.randn(y.shape, generator=g)
    (y * probe).sum().backward()
    out = dict(case=case, conv_weight=mod.conv_offset_aff.weight.detach().clone(), conv_bias=mod.conv_offset_aff.bias.detach().clone(),
               aff_scale_const=float(scale), feat_init=feat_init.detach(), sparse=sparse, guidance=guidance.detach(),
               confidence=confidence.detach(), probe=probe, offset_aff=kept['offset_aff'].detach(), offset=offset.detach(),
               aff=aff.detach(), feat_1=feats[0].detach(), feat_9=feats[8].detach(), feat_out=y.detach(),
               g_feat_init=feat_init.grad, g_guidance=guidance.grad, g_confidence=confidence.grad,
               g_offset_aff=kept['offset_aff'].grad, g_conv_weight=mod.conv_offset_aff.weight.grad, g_conv_bias=mod.conv_offset_aff.bias.grad)
    return out


def main():
    NLSPN = load_reference_nlspn()
    for case in CASES:
        fx = run_case(NLSPN, case)
        path = os.path.join(GOLDEN_DIR, case['name'] + '.pt')
        torch.save(fx, path)
        print('wrote %s (%.0f KB): |y| %.4f  |g_feat_init| %.4f  |g_guidance| %.4f  |g_conf| %.4f' % (
            path, os.path.getsize(path) / 1e3, float(fx['feat_out'].norm()), float(fx['g_feat_init'].norm()),
            float(fx['g_guidance'].norm()), float(fx['g_confidence'].norm())))


if __name__ == '__main__':
    main()
