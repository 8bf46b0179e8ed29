"""Drop-in mirror of the reference's `Transforms` (src/transforms.py:8-668) for the augmentations whose arithmetic runs in
libptta_b200.so: the photometric chain (brightness / contrast / gamma / saturation + image normalisation) and the geometric transforms
listed below.

Same constructor arguments, same `transform(images_arr, intrinsics_arr, padding_modes, interpolation_modes,
random_transform_probability)` call and return convention, and -- the part that keeps runs comparable -- the SAME random draws: every
`torch.rand(n_batch, device=device)` of the reference is made here in the same order (src/transforms.py:229-331, 386-407), so that with
equal generator state both implementations augment every sample identically.  What differs is what happens after the draws: one
reduction + one elementwise kernel for the whole photometric chain (csrc/augment.cuh) instead of ~15 tensor ops and a host sync per
sample and transform (`float(factors[b])`).

Also native: the per-sample rotation (torchvision `functional.rotate` = affine grid + grid_sample; the angle comes from
`np.random.rand`, as in the reference) and resize-and-crop (`functional.resize` + crop) -- with these, every augmentation the shipped
adaptation scripts enable (bash/adapt/adapt_msgchn_*.sh: brightness, contrast, saturation, horizontal flip, rotate 5, resize-and-crop
1.0 .. 1.5) runs in the library; the random crop to a common shape is native as well.  Gamma and hue jitter and the additive noise are native too.  Crop-and-pad and resize-and-pad (constant padding; bilinear reductions WITHOUT anti-aliasing, as the torchvision release the reference
pins computes them) and the random point removal are native.  What stays without a kernel raises NotImplementedError (no silent
fallback): padding modes other than 'constant', interpolation modes other than nearest / bilinear.  `resize_scaling_depth` (the depth tensors
divided by the width ratio of the enlargement) is native."""
import ctypes
import math
import random

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, c_void_p

_FLOAT3 = ctypes.c_float * 3


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class Transforms(object):
    def __init__(self, normalized_image_range=None, random_brightness=[-1, -1], random_contrast=[-1, -1], random_gamma=[-1, -1],
                 random_hue=[-1, -1], random_saturation=[-1, -1], random_noise_type='none', random_noise_spread=-1,
                 random_remove_patch_percent_range=[-1, -1], random_remove_patch_size=[1, 1], random_crop_to_shape=[-1, -1],
                 random_flip_type=['none'], random_rotate_max=0, random_crop_and_pad=[-1, -1], random_resize_and_crop=[-1, -1],
                 random_resize_and_pad=[-1, -1], resize_scaling_depth=False):
        if normalized_image_range is not None and len(normalized_image_range) > 2:       # flat mean..., std... list (src/transforms.py:73-80)
            k = len(normalized_image_range) // 2
            normalized_image_range = [tuple(normalized_image_range[:k]), tuple(normalized_image_range[k:])]
        self.normalized_image_range = normalized_image_range
        self.do_random_brightness = -1 not in random_brightness
        self.random_brightness = random_brightness
        self.do_random_contrast = -1 not in random_contrast
        self.random_contrast = random_contrast
        self.do_random_saturation = -1 not in random_saturation
        self.random_saturation = random_saturation
        self.do_random_gamma = -1 not in random_gamma
        self.random_gamma = random_gamma
        self.do_random_hue = -1 not in random_hue
        self.random_hue = random_hue
        self.do_random_noise = random_noise_type != 'none' and random_noise_spread > -1
        self.random_noise_type, self.random_noise_spread = random_noise_type, random_noise_spread
        if self.do_random_noise and random_noise_type not in ('gaussian', 'uniform'):
            raise ValueError('Unsupported noise type: {}'.format(random_noise_type))
        self.do_random_resize_and_pad = -1 not in random_resize_and_pad
        self.random_resize_and_pad_min, self.random_resize_and_pad_max = random_resize_and_pad[0], random_resize_and_pad[1]
        if self.do_random_resize_and_pad:
            assert self.random_resize_and_pad_min < self.random_resize_and_pad_max
            assert self.random_resize_and_pad_min > 0
            assert self.random_resize_and_pad_max <= 1.0
        self.resize_scaling_depth = resize_scaling_depth                                                        # :183
        # as in the reference, gamma alone does not trigger the uint8 cast (src/transforms.py:74-78 leaves it out of do_photometric_transforms)
        self.do_photometric_transforms = self.do_random_brightness or self.do_random_contrast or self.do_random_hue or self.do_random_saturation
        self.do_image_normalization = normalized_image_range is not None
        self.do_random_horizontal_flip = 'horizontal' in random_flip_type
        self.do_random_vertical_flip = 'vertical' in random_flip_type
        self.do_random_crop_to_shape = -1 not in random_crop_to_shape
        self.random_crop_to_shape = list(random_crop_to_shape)
        if self.do_random_crop_to_shape and len(random_crop_to_shape) not in (2, 4):
            raise ValueError('Unsupported input for random crop to shape: {}'.format(random_crop_to_shape))
        self.do_random_crop_and_pad = -1 not in random_crop_and_pad
        self.random_crop_and_pad_min, self.random_crop_and_pad_max = random_crop_and_pad[0], random_crop_and_pad[1]
        if self.do_random_crop_and_pad:
            assert self.random_crop_and_pad_min < self.random_crop_and_pad_max
            assert self.random_crop_and_pad_max <= 1
        self.do_random_remove_patch = -1 not in random_remove_patch_percent_range
        self.random_remove_patch_percent_range = random_remove_patch_percent_range
        if len(random_remove_patch_size) == 4:          # a range of (odd) sizes to choose from, src/transforms.py:131-137
            self.random_remove_patch_size_height = list(range(random_remove_patch_size[0], random_remove_patch_size[2] + 2, 2))
            self.random_remove_patch_size_width = list(range(random_remove_patch_size[1], random_remove_patch_size[3] + 2, 2))
        else:
            self.random_remove_patch_size_height = [random_remove_patch_size[0]]
            self.random_remove_patch_size_width = [random_remove_patch_size[1]]
        self.do_random_rotate = random_rotate_max > 0
        self.random_rotate_max = random_rotate_max
        self.do_random_resize_and_crop = -1 not in random_resize_and_crop
        self.random_resize_and_crop_min, self.random_resize_and_crop_max = random_resize_and_crop[0], random_resize_and_crop[1]
        if self.do_random_resize_and_crop:
            assert self.random_resize_and_crop_min < self.random_resize_and_crop_max
            assert self.random_resize_and_crop_min >= 1.0
        self._norm = self._normalisation(normalized_image_range)
        self.rand_device = None        # where the draws are made; None = the images' device, as in the reference
        self._streams = {}

    @staticmethod
    def _normalisation(rng):
        """(mode, mean3, std3) of src/transforms.py:669-712"""
        if rng is None or rng == [0, 255]:
            return 0, None, None
        if any(isinstance(v, (tuple, list)) for v in rng):
            return 3, tuple(float(x) for x in rng[0]), tuple(float(x) for x in rng[1])
        if rng == [0, 1]:
            return 1, None, None
        if rng == [-1, 1]:
            return 2, None, None
        raise ValueError('Unsupported normalization range: {}'.format(rng))

    def _rand(self, n, device):
        return torch.rand(n, device=self.rand_device if self.rand_device is not None else device).to(device)

    def _rng_stream(self, device):
        """The draws (tiny tensors, two host reads for the resize sizes) run on a stream of their own: the host then waits for THEM, not
        for whatever the caller's stream still has queued (the previous frame's adaptation step).  The generator is advanced on the host
        at call time, so the values do not depend on the stream."""
        st = self._streams.get(device)
        if st is None:
            st = self._streams[device] = torch.cuda.Stream(device)
        return st

    def transform(self, images_arr, intrinsics_arr=[], padding_modes=['constant'], interpolation_modes=['nearest'],
                  random_transform_probability=0.00):
        images_arr = list(images_arr)
        device = images_arr[0].device
        if device.type != 'cuda':
            raise RuntimeError('the augmentations run on CUDA only (no CPU fallback); got device %s' % device)
        if images_arr[0].ndim != 4:
            raise ValueError('Unsupported number of dimensions: {}'.format(images_arr[0].ndim))
        n_batch, n_channel = images_arr[0].shape[:2]
        n_height, n_width = images_arr[0].shape[-2:]
        rdev = self.rand_device if self.rand_device is not None else device
        cur = torch.cuda.current_stream(device)
        plan = self._draw(n_batch, n_height, n_width, device, rdev, random_transform_probability, [tuple(im.shape) for im in images_arr])
        for t in plan['device_tensors']:
            t.record_stream(cur)                       # allocated on the draw stream, read by kernels on the caller's stream
        cur.wait_stream(self._rng_stream(device))
        # ---- kernels, in the reference's order ----
        if self.do_photometric_transforms or self.do_image_normalization or self.do_random_noise:
            images_arr = [self._photometric(im, plan['flags'], plan['factors'], plan.get('do_n'), plan['noise'][k] if 'noise' in plan else None)
                          for k, im in enumerate(images_arr)]
        else:
            images_arr = [im.float() for im in images_arr]
        if n_channel == 1:
            images_arr = [im[..., 0:1, :, :] for im in images_arr]
        intrinsics_arr = list(intrinsics_arr)
        if plan.get('crop'):                                                                                     # :337-383
            ch, cw, sy, sx = plan['crop']
            images_arr = [self._crop(im, ch, cw, sy, sx) for im in images_arr]
            off = torch.ones(n_batch)
            intrinsics_arr = self._adjust_intrinsics(intrinsics_arr, x_offsets=off * float(n_width - cw), y_offsets=off * float(n_height - ch))
            n_height, n_width = ch, cw
        if plan.get('do_h') is not None or plan.get('do_v') is not None:                                         # :386-404
            images_arr = [self._flip(im, plan.get('do_h'), plan.get('do_v')) for im in images_arr]
        modes = self._modes(interpolation_modes, len(images_arr))
        if 'rotate' in plan:                                                                                     # :406-423
            do_rotate, theta = plan['rotate']
            images_arr = [self._resample('rotate', im, m, do_rotate, theta) for im, m in zip(images_arr, modes)]
        if 'resize' in plan:                                                                                     # :425-502
            do_rs, r_height, r_width, args = plan['resize']
            intrinsics_arr = self._adjust_intrinsics(intrinsics_arr, x_scales=(r_width / n_width), y_scales=(r_height / n_height))
            images_arr = [self._resample('resize_crop', im, m, do_rs, *args) for im, m in zip(images_arr, modes)]
            if self.resize_scaling_depth:                                                                        # :1274-1275: every tensor but the first
                divisor = (r_width / n_width).to(device=images_arr[0].device, dtype=torch.float32)
                for im in images_arr[1:]:
                    check(_lib.lib().ptta_augment_divide_samples(ptr(im), im.shape[0], im[0].numel(), ptr(do_rs), ptr(divisor), _stream()),
                          'augment_divide_samples')
            intrinsics_arr = self._adjust_intrinsics(intrinsics_arr, x_offsets=(r_width - n_width), y_offsets=(r_height - n_height))
        if 'crop_pad' in plan:                                                                                   # :508-566
            if any(m != 'constant' for m in padding_modes):
                raise NotImplementedError('crop-and-pad with padding modes other than constant')
            do_cp, win = plan['crop_pad']
            images_arr = [self._crop_pad(im, do_cp, win) for im in images_arr]
        if 'resize_pad' in plan:                                                                                 # :578-622
            if any(m != 'constant' for m in padding_modes):
                raise NotImplementedError('resize-and-pad with padding modes other than constant')
            do_rp, geo = plan['resize_pad']
            images_arr = [self._resize_pad(im, m, do_rp, geo) for im, m in zip(images_arr, modes)]
        if 'remove' in plan:                                                                                     # :644-652, 878-953
            images_arr = [self._remove_patches(im, plan['remove'], rdev) for im in images_arr]
        outputs = []
        if len(images_arr) > 0:
            outputs.append(images_arr)
        if len(intrinsics_arr) > 0:
            outputs.append(list(intrinsics_arr))
        return outputs[0] if len(outputs) == 1 else outputs

    def _draw(self, n_batch, n_height, n_width, device, rdev, probability, shapes):
        """every random number of one `transform` call, drawn in the reference's order (src/transforms.py:229-480) on the draw stream;
        none of them depends on image data, only on the shapes"""
        plan = {'flags': {}, 'factors': {}, 'device_tensors': []}

        def keep(t):
            t = t.contiguous()
            if t.is_cuda:
                plan['device_tensors'].append(t)
            return t
        with torch.cuda.stream(self._rng_stream(device)):
            do_random_transform = self._rand(n_batch, device) <= probability                                     # :229-230
            for name, enabled, rng, ge in (('b', self.do_random_brightness, self.random_brightness, True),
                                           ('c', self.do_random_contrast, self.random_contrast, False),
                                           ('g', self.do_random_gamma, self.random_gamma, False),
                                           ('h', self.do_random_hue, self.random_hue, False),
                                           ('s', self.do_random_saturation, self.random_saturation, False)):
                if not enabled:
                    continue
                roll = self._rand(n_batch, device)
                # brightness is applied when its roll is >= 0.5 (:243-245), the others when it is <= 0.5 (:256-258, :295-297)
                plan['flags'][name] = keep(torch.logical_and(do_random_transform, roll >= 0.50 if ge else roll <= 0.50).to(torch.uint8))
                values = self._rand(n_batch, device)
                lo, hi = rng
                plan['factors'][name] = keep((hi - lo) * values + lo)
            if self.do_random_noise:                                                                             # :320-331, 839-875
                do_n = torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50)
                plan['do_n'] = keep(do_n.to(torch.uint8))
                host = do_n.tolist()
                plan['noise'] = []
                for shape in shapes:                        # per tensor, per flagged sample: one draw of the sample's shape, in the reference's order
                    nz = torch.zeros(shape, dtype=torch.float32, device=device)
                    for b in range(n_batch):
                        if host[b]:
                            draw = torch.randn(*shape[1:], device=rdev) if self.random_noise_type == 'gaussian' else torch.rand(*shape[1:], device=rdev)
                            nz[b] = draw.to(device)
                    plan['noise'].append(keep(nz))
            if self.do_random_crop_to_shape:                                                                     # :337-366
                # `do and rand(1) <= 0.5 or range`: the roll is drawn in both forms; two numbers = crop to that shape half of the time, four = always
                roll = bool(torch.rand(1, device=rdev) <= 0.50)
                if len(self.random_crop_to_shape) == 2:
                    do_crop = roll
                    ch, cw = self.random_crop_to_shape
                else:
                    do_crop = True
                    ch = int(np.random.randint(low=self.random_crop_to_shape[0], high=self.random_crop_to_shape[2] + 1))
                    cw = int(np.random.randint(low=self.random_crop_to_shape[1], high=self.random_crop_to_shape[3] + 1))
                if do_crop:
                    start_y = torch.randint(low=0, high=n_height - ch + 1, size=(n_batch,), device=rdev)
                    start_x = torch.randint(low=0, high=n_width - cw + 1, size=(n_batch,), device=rdev)
                    plan['crop'] = (ch, cw, keep(start_y.to(device=device, dtype=torch.int32)), keep(start_x.to(device=device, dtype=torch.int32)))
                    n_height, n_width = ch, cw
            if self.do_random_horizontal_flip:                                                                   # :386-394
                plan['do_h'] = keep(torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8))
            if self.do_random_vertical_flip:                                                                     # :396-404
                plan['do_v'] = keep(torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8))
            if self.do_random_rotate:                                                                            # :406-416
                do_rotate = keep(torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8))
                values = np.random.rand(n_batch)                            # the reference draws the angles from numpy's global generator
                angles = (2 * self.random_rotate_max) * values + (-self.random_rotate_max)
                plan['rotate'] = (do_rotate, keep(torch.from_numpy(self._rotation_grid_matrix(angles, n_height, n_width)).to(device)))
            if self.do_random_resize_and_crop:                                                                   # :425-480
                do_rs = keep(torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8))
                r_height = torch.randint(low=int(self.random_resize_and_crop_min * n_height), high=int(self.random_resize_and_crop_max * n_height),
                                         size=(n_batch,), device=rdev)
                r_width = torch.randint(low=int(self.random_resize_and_crop_min * n_width), high=int(self.random_resize_and_crop_max * n_width),
                                        size=(n_batch,), device=rdev)
                rh, rw = r_height.tolist(), r_width.tolist()               # one host read (the reference makes 2 N of them, :463-477)
                start_y, start_x = [], []
                for b in range(n_batch):
                    start_y.append(torch.randint(low=0, high=rh[b] - n_height + 1, size=(1,), device=rdev))
                    start_x.append(torch.randint(low=0, high=rw[b] - n_width + 1, size=(1,), device=rdev))
                start_y, start_x = torch.cat(start_y, dim=0), torch.cat(start_x, dim=0)
                args = [keep(t.to(device=device, dtype=torch.int32)) for t in (r_height, r_width, start_y, start_x)]
                plan['resize'] = (do_rs, keep(r_height), keep(r_width), args)
            if self.do_random_crop_and_pad:                                                                      # :508-557
                do_cp = keep(torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8))
                max_h, min_h = int(self.random_crop_and_pad_max * n_height), int(self.random_crop_and_pad_min * n_height)
                max_w, min_w = int(self.random_crop_and_pad_max * n_width), int(self.random_crop_and_pad_min * n_width)
                rand_h = torch.randint(low=min_h, high=max_h, size=(n_batch,), device=rdev)
                rand_w = torch.randint(low=min_w, high=max_w, size=(n_batch,), device=rdev)
                rh, rw = rand_h.tolist(), rand_w.tolist()
                start_y = torch.cat([torch.randint(low=0, high=max_h - v, size=(1,), device=rdev) for v in rh])
                start_x = torch.cat([torch.randint(low=0, high=max_w - v, size=(1,), device=rdev) for v in rw])
                end_y = torch.minimum(start_y + rand_h, torch.full_like(start_y, n_height))
                end_x = torch.minimum(start_x + rand_w, torch.full_like(start_x, n_width))
                d_h = (n_height - (end_y - start_y)).int()
                pad_top = (d_h * torch.rand(n_batch, device=rdev)).int()
                d_w = (n_width - (end_x - start_x)).int()
                pad_left = (d_w * torch.rand(n_batch, device=rdev)).int()
                win = torch.stack([start_y.int(), start_x.int(), (end_y - start_y).int(), (end_x - start_x).int(), pad_top, pad_left], dim=1)
                plan['crop_pad'] = (do_cp, keep(win.to(device=device, dtype=torch.int32)))
            if self.do_random_resize_and_pad:                                                                    # :578-612
                do_rp = keep(torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8))
                r_height = torch.randint(low=int(self.random_resize_and_pad_min * n_height), high=int(self.random_resize_and_pad_max * n_height),
                                         size=(n_batch,), device=rdev)
                r_width = torch.randint(low=int(self.random_resize_and_pad_min * n_width), high=int(self.random_resize_and_pad_max * n_width),
                                        size=(n_batch,), device=rdev)
                d_h = (n_height - r_height).int()
                pad_top = (d_h * torch.rand(n_batch, device=rdev)).int()
                d_w = (n_width - r_width).int()
                pad_left = (d_w * torch.rand(n_batch, device=rdev)).int()
                pad_top, pad_left = torch.clamp(pad_top, min=0), torch.clamp(pad_left, min=0)
                geo = torch.stack([r_height.int(), r_width.int(), pad_top, pad_left], dim=1)
                plan['resize_pad'] = (do_rp, keep(geo.to(device=device, dtype=torch.int32)))
            if self.do_random_remove_patch:                                                                      # :625-643
                do_rm = torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50)
                values = self._rand(n_batch, device)
                lo, hi = self.random_remove_patch_percent_range
                patch = [[random.choice(self.random_remove_patch_size_height), random.choice(self.random_remove_patch_size_width)]
                         for _ in range(n_batch)]                             # python's global generator, as the reference
                plan['remove'] = (keep(do_rm.to(torch.uint8)), do_rm.tolist(), (hi - lo) * values + lo,
                                  keep(torch.tensor(patch, dtype=torch.int32).to(device)))
        return plan

    # -- geometric helpers -----------------------------------------------------------------------------------
    @staticmethod
    def _modes(interpolation_modes, n):
        """per-tensor interpolation (0 nearest, 1 bilinear) from the reference's names or PIL enums (NEAREST = 0, BILINEAR = 2), the last
        entry repeated for the remaining tensors (src/transforms.py:1056-1059)"""
        out = []
        for m in list(interpolation_modes) + [interpolation_modes[-1]] * max(0, n - len(interpolation_modes)):
            if m in (0, 'nearest'):
                out.append(0)
            elif m in (2, 'bilinear'):
                out.append(1)
            else:
                raise NotImplementedError('interpolation mode %r (nearest and bilinear are native)' % (m,))
        return out[:n]

    @staticmethod
    def _rotation_grid_matrix(angles, h, w):
        """fp32 [N, 6]: what torchvision multiplies its base grid with for `rotate(img, angle)` -- the inverse rotation about the centre
        (double precision, `_get_inverse_affine_matrix` with angle -> -angle, no shear / scale / translation), cast to fp32, transposed and
        divided by (0.5 w, 0.5 h) in fp32 (`_gen_affine_grid`)"""
        out = np.zeros((len(angles), 6), dtype=np.float32)
        half = np.array([0.5 * w, 0.5 * h], dtype=np.float32)
        for b, angle in enumerate(angles):
            rot = math.radians(-float(angle))
            a, bb, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
            theta = np.array([[d, -bb, 0.0], [-c, a, 0.0]], dtype=np.float32)
            resc = theta.T / half                                              # [3, 2]
            out[b, :3], out[b, 3:] = resc[:, 0], resc[:, 1]
        return out

    @staticmethod
    def _crop(images, ch, cw, sy, sx):
        images = images.float().contiguous()
        n, c, h, w = images.shape
        out = torch.empty((n, c, ch, cw), dtype=torch.float32, device=images.device)
        check(_lib.lib().ptta_augment_crop(ptr(images), ptr(out), n, c, h, w, ch, cw, ptr(sy), ptr(sx), _stream()), 'augment_crop')
        return out

    @staticmethod
    def _remove_patches(images, plan, rdev):
        """the selection is data dependent (a random subset of the sample's non-zero pixels, torch.randperm over their count: src/transforms.py
        :926-953), so it is drawn here, after the tensor exists -- in the reference's order: per tensor, per flagged sample"""
        do_rm, host, densities, patch = plan
        images = images.float().contiguous()
        n, c, h, w = images.shape
        sel = torch.zeros((n, h, w), dtype=torch.uint8, device=images.device)
        for b in range(n):
            if not host[b]:
                continue
            ys, xs = (images[b].abs().sum(dim=0) > 0).nonzero(as_tuple=True)
            k = ys.shape[0]
            perm = torch.randperm(k, device=rdev)
            chosen = perm[0:int(densities[b] * k)].to(images.device)
            sel[b, ys[chosen], xs[chosen]] = 1
        out = torch.empty_like(images)
        check(_lib.lib().ptta_augment_remove_patches(ptr(images), ptr(out), n, c, h, w, ptr(do_rm), ptr(sel), ptr(patch), _stream()),
              'augment_remove_patches')
        return out

    @staticmethod
    def _resize_pad(images, mode, do_rp, geo):
        images = images.float().contiguous()
        n, c, h, w = images.shape
        out = torch.empty_like(images)
        check(_lib.lib().ptta_augment_resize_pad(ptr(images), ptr(out), n, c, h, w, ptr(do_rp), ptr(geo), mode, _stream()), 'augment_resize_pad')
        return out

    @staticmethod
    def _crop_pad(images, do_cp, win):
        images = images.float().contiguous()
        n, c, h, w = images.shape
        out = torch.empty_like(images)
        check(_lib.lib().ptta_augment_crop_pad(ptr(images), ptr(out), n, c, h, w, ptr(do_cp), ptr(win), _stream()), 'augment_crop_pad')
        return out

    @staticmethod
    def _resample(kind, images, mode, flags, *args):
        images = images.float().contiguous()
        n, c, h, w = images.shape
        out = torch.empty_like(images)
        L = _lib.lib()
        if kind == 'rotate':
            check(L.ptta_augment_rotate(ptr(images), ptr(out), n, c, h, w, ptr(flags), ptr(args[0].contiguous()), mode, _stream()), 'augment_rotate')
        else:
            check(L.ptta_augment_resize_crop(ptr(images), ptr(out), n, c, h, w, ptr(flags), ptr(args[0]), ptr(args[1]), ptr(args[2]), ptr(args[3]),
                                             mode, _stream()), 'augment_resize_crop')
        return out

    @staticmethod
    def _adjust_intrinsics(intrinsics_arr, x_scales=None, y_scales=None, x_offsets=None, y_offsets=None):
        """src/transforms.py:1330-1378: focal lengths and optical centres scaled, offsets subtracted -- for EVERY sample of the batch, also the
        ones the transform skipped (the reference's behaviour)"""
        out = []
        for K in intrinsics_arr:
            K = K.clone()
            dev = K.device
            xs = torch.ones(len(K), device=dev) if x_scales is None else x_scales.to(dev)
            ys = torch.ones(len(K), device=dev) if y_scales is None else y_scales.to(dev)
            xo = torch.zeros(len(K), device=dev) if x_offsets is None else x_offsets.to(dev)
            yo = torch.zeros(len(K), device=dev) if y_offsets is None else y_offsets.to(dev)
            K[:, 0, 0] = K[:, 0, 0] * xs
            K[:, 0, 2] = K[:, 0, 2] * xs - xo
            K[:, 1, 1] = K[:, 1, 1] * ys
            K[:, 1, 2] = K[:, 1, 2] * ys - yo
            out.append(K)
        return out

    def _photometric(self, images, flags, factors, do_noise=None, noise=None):
        if images.shape[1] != 3:
            if self.do_photometric_transforms:
                raise NotImplementedError('photometric transforms of %d-channel tensors' % images.shape[1])
        images = images.float().contiguous()
        n, c, h, w = images.shape
        if c != 3:                                  # normalisation only (single-channel input to a Transforms without jitter)
            if noise is not None:
                raise NotImplementedError('noise on %d-channel tensors' % c)
            mode, mean, std = self._norm
            if mode == 3:
                raise NotImplementedError('standard normalisation of %d-channel tensors' % c)
            return images / 255.0 if mode == 1 else (2.0 * (images / 255.0) - 1.0 if mode == 2 else images)
        out = torch.empty_like(images)
        mode, mean, std = self._norm
        ws = torch.empty(n, dtype=torch.int64, device=images.device) if 'c' in flags else None
        check(_lib.lib().ptta_augment_photometric(
            ptr(images), ptr(out), n, h, w, ptr(flags.get('b')), ptr(factors.get('b')), ptr(flags.get('c')), ptr(factors.get('c')),
            ptr(flags.get('s')), ptr(factors.get('s')), ptr(flags.get('g')), ptr(factors.get('g')), ptr(flags.get('h')), ptr(factors.get('h')),
            ptr(do_noise), ptr(noise), float(self.random_noise_spread), 1 if self.random_noise_type == 'uniform' else 0,
            1 if self.do_photometric_transforms else 0, mode,
            _FLOAT3(*mean) if mean else None, _FLOAT3(*std) if std else None, ptr(ws), _stream()), 'augment_photometric')
        return out

    @staticmethod
    def _flip(images, do_h, do_v):
        images = images.float().contiguous()
        n, c, h, w = images.shape
        out = torch.empty_like(images)
        check(_lib.lib().ptta_augment_flip(ptr(images), ptr(out), n, c, h, w, ptr(do_h), ptr(do_v), _stream()), 'augment_flip')
        return out

    def map_interpolation_mode_names_to_enums(self, interpolation_mode_names):
        """src/transforms.py:1380-1401 -- kept for the drivers that call it; the native flips do not resample"""
        return list(interpolation_mode_names)
