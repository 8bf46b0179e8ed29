"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/transforms.pt by running the REAL reference's `Transforms` class
(/root/reference/src/transforms.py, CPU tensors) on seeded inputs.  Each case stores the constructor arguments, the seed (inputs and
draws come from it: torch.manual_seed(seed); inputs first, then the class's own torch.rand calls) and the outputs.
    python oracle/gen_golden_transforms.py"""
import importlib
import os
import random
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REFERENCE_SRC = '/root/reference/src'
IMAGENET = [[0.485, 0.456, 0.406], [0.229, 0.224, 0.225]]

CASES = [
    dict(name='photo_all_p1', seed=11, n=4, h=24, w=40, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[0, 1], random_brightness=[0.6, 1.4], random_contrast=[0.6, 1.4], random_saturation=[0.6, 1.4])),
    dict(name='photo_all_p05', seed=12, n=8, h=16, w=48, prob=0.5, kinds=['image'],
         ctor=dict(normalized_image_range=[0, 1], random_brightness=[0.5, 1.5], random_contrast=[0.5, 1.5], random_saturation=[0.5, 1.5])),
    dict(name='photo_sat_imagenet', seed=13, n=3, h=20, w=28, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=IMAGENET, random_saturation=[0.6, 1.4])),
    dict(name='photo_bright_pm1', seed=14, n=5, h=12, w=36, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[-1, 1], random_brightness=[0.6, 1.4])),
    dict(name='photo_gamma', seed=25, n=6, h=16, w=24, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[0, 1], random_brightness=[0.6, 1.4], random_contrast=[0.6, 1.4], random_gamma=[0.6, 1.6],
                   random_saturation=[0.6, 1.4])),
    dict(name='photo_flat_imagenet', seed=26, n=3, h=12, w=20, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[0.485, 0.456, 0.406, 0.229, 0.224, 0.225], random_contrast=[0.6, 1.4])),
    dict(name='photo_hue', seed=27, n=6, h=16, w=24, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[0, 1], random_hue=[-0.1, 0.1])),
    dict(name='photo_all5', seed=28, n=8, h=16, w=24, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[0, 1], random_brightness=[0.6, 1.4], random_contrast=[0.6, 1.4], random_gamma=[0.7, 1.4],
                   random_hue=[-0.5, 0.5], random_saturation=[0.6, 1.4])),
    dict(name='noise_gaussian', seed=29, n=6, h=12, w=20, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[0, 1], random_noise_type='gaussian', random_noise_spread=4.0, random_saturation=[0.8, 1.2])),
    dict(name='noise_uniform', seed=30, n=6, h=12, w=20, prob=1.0, kinds=['image'],
         ctor=dict(normalized_image_range=[0, 255], random_noise_type='uniform', random_noise_spread=10.0)),
    dict(name='norm_only', seed=15, n=2, h=12, w=20, prob=1.0, kinds=['image'], ctor=dict(normalized_image_range=[0, 1])),
    dict(name='flip_hv', seed=16, n=6, h=14, w=22, prob=1.0, kinds=['image', 'depth', 'depth', 'depth'],
         ctor=dict(random_flip_type=['horizontal', 'vertical'])),
    dict(name='rotate5', seed=18, n=6, h=20, w=32, prob=1.0, kinds=['image', 'depth', 'depth', 'depth'], modes=['bilinear', 'nearest', 'nearest', 'nearest'],
         ctor=dict(random_rotate_max=5)),
    dict(name='rotate25_p05', seed=19, n=8, h=16, w=16, prob=0.5, kinds=['image', 'depth'], modes=['bilinear', 'nearest'], ctor=dict(random_rotate_max=25)),
    dict(name='resize_crop', seed=20, n=6, h=16, w=24, prob=1.0, kinds=['image', 'depth', 'depth', 'depth'], modes=['bilinear', 'nearest', 'nearest', 'nearest'],
         intrinsics=True, ctor=dict(random_resize_and_crop=[1.0, 1.5])),
    dict(name='resize_crop_scaled', seed=34, n=6, h=16, w=24, prob=1.0, kinds=['image', 'depth', 'depth'], modes=['bilinear', 'nearest', 'nearest'],
         intrinsics=True, ctor=dict(random_resize_and_crop=[1.0, 1.5], resize_scaling_depth=True)),
    # the geometric set of the shipped adaptation scripts (bash/adapt/adapt_msgchn_vkitti.sh:37-41)
    dict(name='adapt_script_geometric', seed=21, n=8, h=24, w=40, prob=1.0, kinds=['image', 'depth', 'depth', 'depth'],
         modes=['bilinear', 'nearest', 'nearest', 'nearest'], intrinsics=True,
         ctor=dict(random_flip_type=['horizontal'], random_rotate_max=5, random_resize_and_crop=[1.0, 1.5])),
    dict(name='crop_exact', seed=22, n=6, h=20, w=32, prob=1.0, kinds=['image', 'depth', 'depth'], intrinsics=True,
         ctor=dict(random_crop_to_shape=[12, 24], random_flip_type=['horizontal'])),
    dict(name='crop_exact_b', seed=23, n=4, h=20, w=32, prob=1.0, kinds=['image', 'depth'], intrinsics=True, ctor=dict(random_crop_to_shape=[16, 16])),
    dict(name='crop_range_resize', seed=24, n=5, h=24, w=40, prob=1.0, kinds=['image', 'depth', 'depth', 'depth'],
         modes=['bilinear', 'nearest', 'nearest', 'nearest'], intrinsics=True,
         ctor=dict(random_crop_to_shape=[12, 20, 20, 32], random_resize_and_crop=[1.0, 1.5], random_rotate_max=10)),
    dict(name='crop_and_pad', seed=31, n=8, h=20, w=32, prob=1.0, kinds=['image', 'depth', 'depth'],
         ctor=dict(random_crop_and_pad=[0.5, 1.0], random_flip_type=['horizontal'])),
    dict(name='remove_points', seed=32, n=8, h=24, w=40, prob=1.0, kinds=['depth'],
         ctor=dict(random_remove_patch_percent_range=[0.2, 0.6], random_remove_patch_size=[1, 1, 5, 7])),
    # resize-and-pad: the installed torchvision (0.26) low-pass filters a BILINEAR reduction by default, the release the reference pins (0.10.1)
    # has no such option -- the fixture holds what the class produces HERE; the nearest-neighbour tensors are release independent
    dict(name='resize_and_pad', seed=33, n=8, h=24, w=40, prob=1.0, kinds=['image', 'depth', 'depth'], modes=['bilinear', 'nearest', 'nearest'],
         ctor=dict(random_resize_and_pad=[0.6, 1.0])),
    dict(name='flip_h_p05', seed=17, n=8, h=10, w=18, prob=0.5, kinds=['image', 'depth'], ctor=dict(random_flip_type=['horizontal'])),
]


def case_inputs(case):
    """seeded inputs: smooth + textured [0,255] images (so that grey means are not all alike), sparse-ish depth maps"""
    torch.manual_seed(case['seed'])
    n, h, w = case['n'], case['h'], case['w']
    out = []
    for kind in case['kinds']:
        if kind == 'image':
            y = torch.linspace(0, 1, h).view(1, 1, h, 1)
            x = torch.linspace(0, 1, w).view(1, 1, 1, w)
            base = 255 * torch.stack([x.expand(n, 1, h, w), y.expand(n, 1, h, w), (x * y).expand(n, 1, h, w)], 1).squeeze(2)
            scale = torch.rand(n, 3, 1, 1) * 0.8 + 0.2
            out.append((base * scale + 40 * torch.rand(n, 3, h, w)).clamp(0, 255))
        else:
            out.append(torch.rand(n, 1, h, w) * 80 * (torch.rand(n, 1, h, w) < 0.3).float())
    return out


def case_intrinsics(case):
    K = torch.eye(3).repeat(case['n'], 1, 1)
    K[:, 0, 0] = 700.0 + torch.arange(case['n']); K[:, 1, 1] = 710.0
    K[:, 0, 2] = case['w'] / 2.0; K[:, 1, 2] = case['h'] / 2.0
    return K


def main():
    if not hasattr(Image, 'ANTIALIAS'):
        Image.ANTIALIAS = Image.LANCZOS          # the reference's constructor names the pre-Pillow-10 constant (src/transforms.py:187)
    sys.path.insert(0, REFERENCE_SRC)
    T = importlib.import_module('transforms')
    fixtures = {}
    for case in CASES:
        inputs = case_inputs(case)
        np.random.seed(case['seed'])
        random.seed(case['seed'])
        tr = T.Transforms(**case['ctor'])
        kw = {}
        if 'modes' in case:
            kw['interpolation_modes'] = tr.map_interpolation_mode_names_to_enums(case['modes'])
        Ks = None
        if case.get('intrinsics'):
            kw['intrinsics_arr'] = [case_intrinsics(case)]
        outs = tr.transform(images_arr=[t.clone() for t in inputs], random_transform_probability=case['prob'], **kw)
        if case.get('intrinsics'):
            outs, Ks = outs
        fixtures[case['name']] = {'case': case, 'outputs': [o.clone() for o in outs], 'intrinsics': [k.clone() for k in Ks] if Ks else None}
        print('%-20s %s' % (case['name'], ' '.join('%.6f' % float(o.mean()) for o in outs)))
    path = os.path.join(ROOT, 'tests', 'golden', 'transforms.pt')
    torch.save({'fixtures': fixtures, 'torch_version': torch.__version__}, path)
    print('%s  %.0f KB' % (os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
