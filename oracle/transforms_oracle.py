"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's `Transforms.transform` (src/transforms.py:192-668) for the options
the native path implements: brightness / contrast / saturation jitter (:236-301, per-sample torchvision tensor ops, :714-837), image
normalisation (:669-712), random crop to a common shape (:337-383, 955-988), per-sample flips (:386-407, 990-1034), rotation (:406-423, 1036-1070) and resize-and-crop (:425-502, 1222-1283).  The third-party arithmetic is torchvision's
(`torchvision.transforms.functional.adjust_*`, pinned 0.10.1 by the reference's README.md:74; 0.26 here -- the tensor code path
`_blend` / `rgb_to_grayscale` is unchanged between the two) and is called as the reference calls it.  Pinned against fixtures produced by
the reference's own class: oracle/gen_golden_transforms.py, tests/test_transforms_oracle.py."""
import random

import numpy as np
import torch
from torchvision.transforms import functional
from torchvision.transforms import InterpolationMode

_MODE = {'nearest': InterpolationMode.NEAREST, 'bilinear': InterpolationMode.BILINEAR, 0: InterpolationMode.NEAREST, 2: InterpolationMode.BILINEAR}


def draws(n_batch, cfg, probability, rand=lambda n: torch.rand(n)):
    """the reference's random draws, in its order (T:229-331, 386-404); cfg: dict with optional 'brightness' / 'contrast' /
    'saturation' ranges and 'flip' = subset of ('horizontal', 'vertical')"""
    d = {'do': rand(n_batch) <= probability}
    for name, ge in (('brightness', True), ('contrast', False), ('gamma', False), ('hue', False), ('saturation', False)):       # T:242-301
        if name in cfg:
            roll = rand(n_batch)
            d['do_' + name] = torch.logical_and(d['do'], roll >= 0.50 if ge else roll <= 0.50)
            lo, hi = cfg[name]
            d['f_' + name] = (hi - lo) * rand(n_batch) + lo
    if 'noise' in cfg:                                                          # T:320-331, 839-875: one draw per tensor and flagged sample
        kind, spread = cfg['noise']
        d['do_noise'] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
        d['noise'] = []
        for shape in cfg['tensor_shapes']:
            per = {}
            for b in range(n_batch):
                if d['do_noise'][b]:
                    per[b] = torch.randn(*shape[1:]) if kind == 'gaussian' else torch.rand(*shape[1:])
            d['noise'].append(per)
    if 'crop_to_shape' in cfg:                                                  # T:337-366
        spec = cfg['crop_to_shape']
        h, w = cfg['shape']
        roll = bool(torch.rand(1) <= 0.50)
        if len(spec) == 2:
            d['do_crop'], (d['crop_h'], d['crop_w']) = roll, spec
        else:
            d['do_crop'] = True
            d['crop_h'] = int(np.random.randint(low=spec[0], high=spec[2] + 1))
            d['crop_w'] = int(np.random.randint(low=spec[1], high=spec[3] + 1))
        if d['do_crop']:
            d['crop_y'] = torch.randint(low=0, high=h - d['crop_h'] + 1, size=(n_batch,))
            d['crop_x'] = torch.randint(low=0, high=w - d['crop_w'] + 1, size=(n_batch,))
            cfg = dict(cfg, shape=(d['crop_h'], d['crop_w']))                   # later transforms see the cropped size (T:381-383)
    for name in ('horizontal', 'vertical'):
        if name in cfg.get('flip', ()):
            d['do_' + name] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
    if 'rotate' in cfg:                                                         # T:406-416 (angles from numpy's global generator)
        d['do_rotate'] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
        m = cfg['rotate']
        d['angles'] = (m - (-m)) * np.random.rand(n_batch) + (-m)
    if 'resize_and_crop' in cfg:                                                # T:425-480
        lo, hi = cfg['resize_and_crop']
        h, w = cfg['shape']
        d['do_resize_and_crop'] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
        d['r_height'] = torch.randint(low=int(lo * h), high=int(hi * h), size=(n_batch,))
        d['r_width'] = torch.randint(low=int(lo * w), high=int(hi * w), size=(n_batch,))
        sy, sx = [], []
        for b in range(n_batch):
            sy.append(torch.randint(low=0, high=int(d['r_height'][b]) - h + 1, size=(1,)))
            sx.append(torch.randint(low=0, high=int(d['r_width'][b]) - w + 1, size=(1,)))
        d['start_y'], d['start_x'] = torch.cat(sy), torch.cat(sx)
    if 'crop_and_pad' in cfg:                                                   # T:508-557
        lo, hi = cfg['crop_and_pad']
        h, w = cfg['shape']
        d['do_crop_and_pad'] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
        max_h, min_h, max_w, min_w = int(hi * h), int(lo * h), int(hi * w), int(lo * w)
        rand_h = torch.randint(low=min_h, high=max_h, size=(n_batch,))
        rand_w = torch.randint(low=min_w, high=max_w, size=(n_batch,))
        sy = torch.cat([torch.randint(low=0, high=max_h - int(v), size=(1,)) for v in rand_h])
        sx = torch.cat([torch.randint(low=0, high=max_w - int(v), size=(1,)) for v in rand_w])
        ey = torch.minimum(sy + rand_h, torch.full_like(sy, h))
        ex = torch.minimum(sx + rand_w, torch.full_like(sx, w))
        dh = (h - (ey - sy)).int()
        pt = (dh * torch.rand(n_batch)).int()
        dw = (w - (ex - sx)).int()
        pl = (dw * torch.rand(n_batch)).int()
        d['cp'] = (sy, sx, ey, ex, pt, dh - pt, pl, dw - pl)
    if 'resize_and_pad' in cfg:                                                 # T:578-612
        lo, hi = cfg['resize_and_pad']
        h, w = cfg['shape']
        d['do_resize_and_pad'] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
        rh = torch.randint(low=int(lo * h), high=int(hi * h), size=(n_batch,))
        rw = torch.randint(low=int(lo * w), high=int(hi * w), size=(n_batch,))
        dh = (h - rh).int()
        pt = (dh * torch.rand(n_batch)).int()
        dw = (w - rw).int()
        pl = (dw * torch.rand(n_batch)).int()
        zero = torch.zeros_like(pt)
        d['rp'] = (rh, rw, torch.maximum(pt, zero), torch.maximum(dh - pt, zero), torch.maximum(pl, zero), torch.maximum(dw - pl, zero))
    if 'remove_patch' in cfg:                                                   # T:625-643 (patch sizes from python's global generator)
        (lo, hi), heights, widths = cfg['remove_patch']
        d['do_remove'] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
        d['densities'] = (hi - lo) * rand(n_batch) + lo
        d['patch'] = [[random.choice(heights), random.choice(widths)] for _ in range(n_batch)]
    return d


def apply(images_arr, cfg, d, normalized_image_range=None, interpolation_modes=('nearest',), antialias=False):
    images_arr = [im.clone() for im in images_arr]
    photometric = any(k in cfg for k in ('brightness', 'contrast', 'hue', 'saturation'))        # gamma alone does not trigger the cast (T:102-106)
    if photometric:
        images_arr = [im.to(torch.uint8) if torch.is_floating_point(im) else im for im in images_arr]          # T:236-240
    for name, fn in (('brightness', functional.adjust_brightness), ('contrast', functional.adjust_contrast), ('gamma', functional.adjust_gamma),
                     ('hue', functional.adjust_hue), ('saturation', functional.adjust_saturation)):
        if name in cfg:
            for images in images_arr:
                for b in range(images.shape[0]):
                    if d['do_' + name][b]:
                        images[b, ...] = fn(images[b], d['f_' + name][b])                                        # T:714-837
    images_arr = [im.float() for im in images_arr]
    if 'noise' in cfg:
        kind, spread = cfg['noise']
        for images, per in zip(images_arr, d['noise']):
            for b, nz in per.items():
                images[b, ...] = images[b] + spread * nz if kind == 'gaussian' else images[b] + spread * (nz - 0.5)
    rng = normalized_image_range
    if rng is not None:                                                                                            # T:669-712
        if rng == [0, 1]:
            images_arr = [im / 255.0 for im in images_arr]
        elif any(isinstance(v, (tuple, list)) for v in rng):
            images_arr = [functional.normalize(im / 255.0, rng[0], rng[1]) for im in images_arr]
        elif rng == [-1, 1]:
            images_arr = [2.0 * (im / 255.0) - 1.0 for im in images_arr]
        elif rng != [0, 255]:
            raise ValueError(rng)
    if d.get('do_crop'):                                                        # T:955-988
        ch, cw = d['crop_h'], d['crop_w']
        images_arr = [torch.stack([im[b, :, int(d['crop_y'][b]):int(d['crop_y'][b]) + ch, int(d['crop_x'][b]):int(d['crop_x'][b]) + cw]
                                   for b in range(im.shape[0])], dim=0) for im in images_arr]
    for name, dim in (('horizontal', -1), ('vertical', -2)):
        if 'do_' + name in d:
            for images in images_arr:
                for b in range(images.shape[0]):
                    if d['do_' + name][b]:
                        images[b, ...] = torch.flip(images[b], dims=[dim])                                       # T:990-1034
    if 'do_rotate' in d:                                                        # T:1036-1070
        modes = list(interpolation_modes) + [interpolation_modes[-1]] * (len(images_arr) - len(interpolation_modes))
        for images, mode in zip(images_arr, modes):
            for b in range(images.shape[0]):
                if d['do_rotate'][b]:
                    images[b, ...] = functional.rotate(images[b], angle=d['angles'][b], interpolation=_MODE[mode], expand=False)
    if 'do_resize_and_crop' in d:                                               # T:1222-1283
        modes = list(interpolation_modes) + [interpolation_modes[-1]] * (len(images_arr) - len(interpolation_modes))
        h, w = images_arr[0].shape[-2:]
        for i, (images, mode) in enumerate(zip(images_arr, modes)):
            out = []
            for b in range(images.shape[0]):
                image = images[b]
                if d['do_resize_and_crop'][b]:
                    image = functional.resize(image, size=(int(d['r_height'][b]), int(d['r_width'][b])), interpolation=_MODE[mode])
                    y0, x0 = int(d['start_y'][b]), int(d['start_x'][b])
                    image = image[..., y0:y0 + h, x0:x0 + w]
                    if i != 0 and cfg.get('resize_scaling_depth'):              # T:1274-1275: int64 0-d tensor / int -> float32 0-d tensor
                        image = image / (d['r_width'][b] / w)
                out.append(image)
            images_arr[i] = torch.stack(out, dim=0)
    if 'do_crop_and_pad' in d:                                                  # T:1072-1135 (constant padding)
        sy, sx, ey, ex, pt, pb, pl, pr = d['cp']
        for images in images_arr:
            for b in range(images.shape[0]):
                if d['do_crop_and_pad'][b]:
                    image = images[b][..., int(sy[b]):int(ey[b]), int(sx[b]):int(ex[b])]
                    images[b, ...] = functional.pad(image, (int(pl[b]), int(pt[b]), int(pr[b]), int(pb[b])), padding_mode='constant', fill=0)
    if 'do_resize_and_pad' in d:                                                # T:1137-1220; antialias=False = the pinned torchvision 0.10.1 (no such option there)
        modes = list(interpolation_modes) + [interpolation_modes[-1]] * (len(images_arr) - len(interpolation_modes))
        rh, rw, pt, pb, pl, pr = d['rp']
        for images, mode in zip(images_arr, modes):
            for b in range(images.shape[0]):
                if d['do_resize_and_pad'][b]:
                    image = functional.resize(images[b], size=(int(rh[b]), int(rw[b])), interpolation=_MODE[mode], antialias=antialias)
                    images[b, ...] = functional.pad(image, (int(pl[b]), int(pt[b]), int(pr[b]), int(pb[b])), padding_mode='constant', fill=0)
    if 'do_remove' in d:                                                        # T:878-953 (the subset is drawn here: it depends on the data)
        for images in images_arr:
            for b in range(images.shape[0]):
                if d['do_remove'][b]:
                    image = images[b]
                    mask = torch.sum(torch.abs(image), dim=0, keepdim=True)
                    mask = torch.where(mask > 0, torch.ones_like(mask), torch.zeros_like(mask))
                    nz = (mask > 0).nonzero(as_tuple=True)
                    subset = torch.randperm(nz[0].shape[0])
                    subset = subset[0:int(d['densities'][b] * subset.shape[0])]
                    mask[tuple(idx[subset] for idx in nz)] = float('inf')
                    ps = d['patch'][b]
                    mask = torch.nn.functional.max_pool2d(input=mask, kernel_size=ps, stride=1, padding=[int(k // 2) for k in ps])
                    mask[mask == float('inf')] = 0.0
                    images[b, ...] = mask * image
    return images_arr


def adjust_intrinsics(intrinsics_arr, d, shape):
    """T:449-453, 498-502 with T:1330-1378: every sample's intrinsics are rescaled and shifted, also those the transform skipped"""
    h, w = shape
    out = []
    for K in intrinsics_arr:
        K = K.clone()
        for b in range(len(K)):
            if d.get('do_crop'):                                                # T:372-379: the full size difference, for every sample
                K[b, 0, 2] = K[b, 0, 2] * 1.0 - float(w - d['crop_w'])
                K[b, 1, 2] = K[b, 1, 2] * 1.0 - float(h - d['crop_h'])
        if d.get('do_crop'):
            h, w = d['crop_h'], d['crop_w']
        for b in range(len(K)):
            if 'do_resize_and_crop' not in d:
                continue
            xs, ys = d['r_width'][b] / w, d['r_height'][b] / h
            K[b, 0, 0] = K[b, 0, 0] * xs
            K[b, 0, 2] = K[b, 0, 2] * xs
            K[b, 1, 1] = K[b, 1, 1] * ys
            K[b, 1, 2] = K[b, 1, 2] * ys
            K[b, 0, 2] = K[b, 0, 2] * 1.0 - (d['r_width'][b] - w)
            K[b, 1, 2] = K[b, 1, 2] * 1.0 - (d['r_height'][b] - h)
        out.append(K)
    return out
