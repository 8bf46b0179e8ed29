"""GPU: the native augmentations (csrc/augment.cuh through tta_depth_completion_b200.transforms.Transforms, the drop-in mirror of
src/transforms.py) against (a) the fixtures produced by the reference's own class and (b) the CPU oracle at the benchmark frame size.
The draws are made on the CPU generator here (`rand_device`), so that both sides see the same random numbers.

Stated tolerance: flips, crops, brightness, hue, saturation, additive noise and every normalisation are BIT-EXACT.  The contrast transform blends with the mean of
the grey image: the native path sums the grey values exactly (integers), torch.mean accumulates in fp32 in an order of its own -- the two
means differ by ~1e-7 relative, which moves a pixel by one uint8 step only when the blended value lands within ~1e-5 of an integer:
at most 1e-4 of the values may differ, each by exactly one grey level.  Rotation / resize-and-crop resample in fp32: see
`compare_resampled`."""
import os

import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from golden_util import GOLDEN_DIR
from oracle import transforms_oracle as TO
from oracle.gen_golden_transforms import case_inputs, IMAGENET
from test_transforms_oracle import FIX, case_cfg, nested_range

DEV = 'cuda'


def level(rng):
    """size of one uint8 step after normalisation"""
    if rng == [0, 1]:
        return 1 / 255.0
    if rng == [-1, 1]:
        return 2 / 255.0
    if rng is None or rng == [0, 255]:
        return 1.0
    return 1 / 255.0 / min(rng[1])


def compare(got, want, contrast, rng, what, max_frac=1e-4, max_levels=1):
    got = got.cpu()
    assert got.shape == want.shape and got.dtype == want.dtype, what
    if not contrast:
        assert torch.equal(got, want), (what, float((got - want).abs().max()))
        return 0.0
    diff = (got - want).abs()
    frac = float((diff > 0).float().mean())
    assert frac <= max_frac and float(diff.max()) <= max_levels * level(rng) * 1.0001 + 1e-6, (what, frac, float(diff.max()))
    return frac


@pytest.mark.parametrize('name', sorted(FIX))
def test_native_transforms_match_reference_fixture(name):
    from tta_depth_completion_b200.transforms import Transforms
    fx = FIX[name]
    case = fx['case']
    inputs = case_inputs(case)
    tr = Transforms(**case['ctor'])
    tr.rand_device = 'cpu'
    np.random.seed(case['seed'])
    random.seed(case['seed'])
    kw = {}
    if 'modes' in case:
        kw['interpolation_modes'] = tr.map_interpolation_mode_names_to_enums(case['modes'])
    if case.get('intrinsics'):
        from oracle.gen_golden_transforms import case_intrinsics
        kw['intrinsics_arr'] = [case_intrinsics(case).to(DEV)]
    outs = tr.transform(images_arr=[t.to(DEV) for t in inputs], random_transform_probability=case['prob'], **kw)
    if case.get('intrinsics'):
        outs, Ks = outs
        assert torch.equal(Ks[0].cpu(), fx['intrinsics'][0])
    assert len(outs) == len(fx['outputs'])
    resampled = 'random_rotate_max' in case['ctor'] or 'random_resize_and_crop' in case['ctor'] or 'random_resize_and_pad' in case['ctor']
    want_all = fx['outputs']
    if 'random_resize_and_pad' in case['ctor']:
        # the native bilinear reduction is the pinned release's (no anti-aliasing); the fixture's image tensor was filtered by the installed one:
        # that tensor is compared with the oracle run with antialias=False, the nearest-neighbour tensors with the fixture itself
        np.random.seed(case['seed']); random.seed(case['seed'])
        inputs2 = case_inputs(case)
        cfg = case_cfg(case)
        plain = TO.apply(inputs2, cfg, TO.draws(case['n'], cfg, case['prob']), None, case['modes'], antialias=False)
        want_all = [plain[0]] + list(fx['outputs'][1:])
        for a, b in zip(plain[1:], fx['outputs'][1:]):
            assert torch.equal(a, b)
    for k, (got, want) in enumerate(zip(outs, want_all)):
        if resampled:
            compare_resampled(got, want, '%s[%d]' % (name, k))
        else:
            # contrast (grey mean) and gamma (powf of the CPU fixture vs the device's) may move a value by one grey level
            soft = 'random_contrast' in case['ctor'] or 'random_gamma' in case['ctor']
            # a one-level difference that enters the hue transform (RGB -> HSV -> RGB) can leave it as up to three levels
            compare(got, want, soft, nested_range(case['ctor'].get('normalized_image_range')), '%s[%d]' % (name, k),
                    max_frac=2e-3 if 'random_gamma' in case['ctor'] else 1e-4, max_levels=3 if (soft and 'random_hue' in case['ctor']) else 1)


def compare_resampled(got, want, what, max_bad=2e-3):
    """rotation / resize: the source coordinate of every output pixel is a chain of fp32 operations that torch evaluates through a batched
    matrix product (rotation) or in another association (resize); a coordinate that differs in its last bit picks another source pixel
    only when it sits on a rounding boundary (nearest) or moves a bilinear weight by ~1e-6.  Stated tolerance: values equal to 1e-4 of the
    value range everywhere except at most 0.2 % of the pixels (the boundary cases)."""
    got = got.cpu()
    assert got.shape == want.shape and got.dtype == want.dtype, what
    scale = max(float(want.abs().max()), 1e-6)
    bad = float(((got - want).abs() > 1e-4 * scale).float().mean())
    assert bad <= max_bad, (what, bad)
    return bad


@pytest.mark.parametrize('n,h,w', [(4, 240, 1216), (1, 352, 1216)])
def test_native_geometric_at_frame_size(n, h, w):
    """the geometric set of the shipped adaptation scripts (horizontal flip, rotate 5, resize-and-crop 1.0 .. 1.5; image bilinear, sparse depth /
    validity / ground truth nearest) on full frames, against the CPU oracle (torchvision on the same draws)"""
    from tta_depth_completion_b200.transforms import Transforms
    from tta_depth_completion_b200.synthetic import synthetic_frame
    image, sparse, dense = synthetic_frame(5, 2, n, h, w, 'kitti')
    validity = (sparse > 0).float()
    cfg = {'flip': ('horizontal',), 'rotate': 5, 'resize_and_crop': [1.0, 1.5], 'shape': (h, w)}
    modes = ['bilinear', 'nearest', 'nearest', 'nearest']
    worst = 0.0
    for seed in range(3):
        torch.manual_seed(200 + seed); np.random.seed(200 + seed)
        d = TO.draws(n, cfg, 1.0)
        want = TO.apply([image, sparse, validity, dense], cfg, d, None, modes)
        tr = Transforms(random_flip_type=['horizontal'], random_rotate_max=5, random_resize_and_crop=[1.0, 1.5])
        tr.rand_device = 'cpu'
        torch.manual_seed(200 + seed); np.random.seed(200 + seed)
        got = tr.transform(images_arr=[t.to(DEV) for t in (image, sparse, validity, dense)], interpolation_modes=modes, random_transform_probability=1.0)
        for k, (g, w_) in enumerate(zip(got, want)):
            worst = max(worst, compare_resampled(g, w_, 'seed %d tensor %d' % (seed, k)))
    print('fraction of pixels beyond 1e-4 of the range (rounding-boundary source pixels): %.2e' % worst)


@pytest.mark.parametrize('n,h,w', [(1, 352, 1216), (4, 240, 1216), (2, 480, 640)])
@pytest.mark.parametrize('rng', [[0, 1], IMAGENET])
def test_native_photometric_at_frame_size(n, h, w, rng):
    """the shipped adaptation scripts' jitter (bash/adapt/adapt_msgchn_vkitti.sh: brightness / contrast / saturation 0.6 .. 1.4,
    probability 1) on full frames, against the CPU oracle on the same draws"""
    from tta_depth_completion_b200.transforms import Transforms
    from tta_depth_completion_b200.synthetic import synthetic_frame
    image = synthetic_frame(3, 0, n, h, w, 'kitti')[0]
    ctor = dict(normalized_image_range=rng, random_brightness=[0.6, 1.4], random_contrast=[0.6, 1.4], random_saturation=[0.6, 1.4])
    cfg = {'brightness': [0.6, 1.4], 'contrast': [0.6, 1.4], 'saturation': [0.6, 1.4]}
    worst = 0.0
    for seed in range(3):
        torch.manual_seed(100 + seed)
        d = TO.draws(n, cfg, 1.0)
        want = TO.apply([image], cfg, d, rng)[0]
        tr = Transforms(**ctor)
        tr.rand_device = 'cpu'
        torch.manual_seed(100 + seed)
        got = tr.transform(images_arr=[image.to(DEV)], random_transform_probability=1.0)[0]
        worst = max(worst, compare(got, want, True, rng, 'seed %d' % seed))
    print('fraction of values off by one grey level (contrast mean): %.2e' % worst)


@pytest.mark.parametrize('n,h,w', [(3, 15, 21), (2, 16, 24), (5, 9, 12)])
def test_scalar_and_vector_paths(n, h, w):
    """odd sizes take the 4-byte path of the photometric / flip kernels, multiples of four the 16-byte path: both against the oracle"""
    from tta_depth_completion_b200.transforms import Transforms
    torch.manual_seed(h * w)
    image = torch.rand(n, 3, h, w) * 255
    depth = torch.rand(n, 1, h, w)
    cfg = {'brightness': [0.5, 1.5], 'contrast': [0.5, 1.5], 'saturation': [0.5, 1.5]}
    torch.manual_seed(5)
    want = TO.apply([image], cfg, TO.draws(n, cfg, 1.0), [0, 1])[0]
    tr = Transforms(normalized_image_range=[0, 1], random_brightness=[0.5, 1.5], random_contrast=[0.5, 1.5], random_saturation=[0.5, 1.5])
    tr.rand_device = 'cpu'
    torch.manual_seed(5)
    got = tr.transform(images_arr=[image.to(DEV)], random_transform_probability=1.0)[0]
    compare(got, want, True, [0, 1], 'photometric %dx%d' % (h, w))
    cfg = {'flip': ('horizontal', 'vertical')}
    torch.manual_seed(6)
    want = TO.apply([image, depth], cfg, TO.draws(n, cfg, 1.0), None)
    tr = Transforms(random_flip_type=['horizontal', 'vertical'])
    tr.rand_device = 'cpu'
    torch.manual_seed(6)
    got = tr.transform(images_arr=[image.to(DEV), depth.to(DEV)], random_transform_probability=1.0)
    for g, w_ in zip(got, want):
        assert torch.equal(g.cpu(), w_)


def test_native_gamma_matches_torchvision_on_the_device():
    """gamma jitter against torchvision running on the SAME device (the reference's setting: src/tta_main.py keeps the tensors on the GPU):
    both evaluate powf with the device's math library"""
    from torchvision.transforms import functional
    from tta_depth_completion_b200.transforms import Transforms
    torch.manual_seed(3)
    n, h, w = 4, 64, 96
    image = (torch.rand(n, 3, h, w) * 255).to(DEV)
    tr = Transforms(normalized_image_range=[0, 255], random_brightness=[0.9, 1.1], random_gamma=[0.5, 2.0])
    tr.rand_device = 'cpu'
    torch.manual_seed(9)
    got = tr.transform(images_arr=[image], random_transform_probability=1.0)[0]
    torch.manual_seed(9)
    d = TO.draws(n, {'brightness': [0.9, 1.1], 'gamma': [0.5, 2.0]}, 1.0)
    want = image.to(torch.uint8)
    for b in range(n):
        if d['do_brightness'][b]:
            want[b] = functional.adjust_brightness(want[b], float(d['f_brightness'][b]))
        if d['do_gamma'][b]:
            want[b] = functional.adjust_gamma(want[b], d['f_gamma'][b].to(DEV))
    diff = (got - want.float()).abs()
    assert float(diff.max()) <= 1.0 and float((diff > 0).float().mean()) <= 1e-4, (float(diff.max()), float((diff > 0).float().mean()))
    assert bool(d['do_gamma'].any())


def test_native_flips_at_frame_size():
    from tta_depth_completion_b200.transforms import Transforms
    from tta_depth_completion_b200.synthetic import synthetic_frame
    image, sparse, dense = synthetic_frame(4, 1, 4, 240, 1216, 'kitti')
    cfg = {'flip': ('horizontal', 'vertical')}
    torch.manual_seed(7)
    d = TO.draws(4, cfg, 1.0)
    want = TO.apply([image, sparse, dense], cfg, d, None)
    tr = Transforms(random_flip_type=['horizontal', 'vertical'])
    tr.rand_device = 'cpu'
    torch.manual_seed(7)
    got = tr.transform(images_arr=[image.to(DEV), sparse.to(DEV), dense.to(DEV)], random_transform_probability=1.0)
    for g, w_ in zip(got, want):
        assert torch.equal(g.cpu(), w_)


def test_unsupported_options_fail_loudly():
    from tta_depth_completion_b200.transforms import Transforms
    with pytest.raises(NotImplementedError, match='padding modes'):
        t = Transforms(random_crop_and_pad=[0.5, 1.0])
        t.rand_device = 'cpu'
        t.transform(images_arr=[torch.zeros(2, 3, 8, 8, device=DEV)], padding_modes=['reflect'], random_transform_probability=1.0)
