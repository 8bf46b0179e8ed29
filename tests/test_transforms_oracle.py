"""CPU: oracle/transforms_oracle.py (restatement of the reference's Transforms for the natively implemented options) against the
fixtures produced by the reference's own class (oracle/gen_golden_transforms.py) -- bit-exact: same torch / torchvision CPU ops, same draws."""
import os

import random

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from oracle import transforms_oracle as TO
from oracle.gen_golden_transforms import case_inputs

FIX = torch.load(os.path.join(GOLDEN_DIR, 'transforms.pt'), weights_only=False)['fixtures']


def nested_range(rng):
    """the constructor's own re-shaping of a flat mean..., std... list (src/transforms.py:73-80)"""
    if rng is not None and len(rng) > 2:
        k = len(rng) // 2
        return [tuple(rng[:k]), tuple(rng[k:])]
    return rng


def case_cfg(case):
    c = case['ctor']
    cfg = {}
    if c.get('random_noise_type', 'none') != 'none':
        cfg['noise'] = (c['random_noise_type'], c['random_noise_spread'])
        cfg['tensor_shapes'] = [(case['n'], 3 if kind == 'image' else 1, case['h'], case['w']) for kind in case['kinds']]
    for k in ('brightness', 'contrast', 'gamma', 'hue', 'saturation'):
        if 'random_' + k in c:
            cfg[k] = c['random_' + k]
    if 'random_flip_type' in c:
        cfg['flip'] = tuple(c['random_flip_type'])
    if 'random_crop_to_shape' in c:
        cfg['crop_to_shape'] = c['random_crop_to_shape']
        cfg['shape'] = (case['h'], case['w'])
    if 'random_crop_and_pad' in c:
        cfg['crop_and_pad'] = c['random_crop_and_pad']
        cfg.setdefault('shape', (case['h'], case['w']))
    if 'random_remove_patch_percent_range' in c:
        sz = c.get('random_remove_patch_size', [1, 1])
        hs = list(range(sz[0], sz[2] + 2, 2)) if len(sz) == 4 else [sz[0]]
        ws = list(range(sz[1], sz[3] + 2, 2)) if len(sz) == 4 else [sz[1]]
        cfg['remove_patch'] = (c['random_remove_patch_percent_range'], hs, ws)
    if 'random_resize_and_pad' in c:
        cfg['resize_and_pad'] = c['random_resize_and_pad']
        cfg.setdefault('shape', (case['h'], case['w']))
    if c.get('random_rotate_max', 0) > 0:
        cfg['rotate'] = c['random_rotate_max']
    if 'random_resize_and_crop' in c:
        cfg['resize_and_crop'] = c['random_resize_and_crop']
        cfg['shape'] = (case['h'], case['w'])
        if c.get('resize_scaling_depth'):
            cfg['resize_scaling_depth'] = True
    return cfg


@pytest.mark.parametrize('name', sorted(FIX))
def test_oracle_equals_reference_transforms(name):
    fx = FIX[name]
    case = fx['case']
    inputs = case_inputs(case)                      # seeds the generator; the draws continue from there, as in the generator script
    cfg = case_cfg(case)
    np.random.seed(case['seed'])
    random.seed(case['seed'])
    d = TO.draws(case['n'], cfg, case['prob'])
    # the fixtures hold what the reference's class produces under the INSTALLED torchvision: anti-aliased bilinear reductions (resize-and-pad)
    outs = TO.apply(inputs, cfg, d, nested_range(case['ctor'].get('normalized_image_range')), case.get('modes', ('nearest',)), antialias=True)
    assert len(outs) == len(fx['outputs'])
    for got, want in zip(outs, fx['outputs']):
        assert got.dtype == want.dtype and torch.equal(got, want)
    if fx.get('intrinsics'):
        from oracle.gen_golden_transforms import case_intrinsics
        Ks = TO.adjust_intrinsics([case_intrinsics(case)], d, (case['h'], case['w']))
        assert torch.equal(Ks[0], fx['intrinsics'][0])
