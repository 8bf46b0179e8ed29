"""Drop-in mirror of the reference's `Transforms` (src/transforms.py:8-668) for the augmentations whose arithmetic runs in
libptta_b200.so: the photometric chain (brightness / contrast / saturation + image normalisation) and the per-sample flips.

Same constructor arguments, same `transform(images_arr, intrinsics_arr, padding_modes, interpolation_modes,
random_transform_probability)` call and return convention, and -- the part that keeps runs comparable -- the SAME random draws: every
`torch.rand(n_batch, device=device)` of the reference is made here in the same order (src/transforms.py:229-331, 386-407), so that with
equal generator state both implementations augment every sample identically.  What differs is what happens after the draws: one
reduction + one elementwise kernel for the whole photometric chain (csrc/augment.cuh) instead of ~15 tensor ops and a host sync per
sample and transform (`float(factors[b])`).

Options that resample the image (rotate, resize-and-crop / -pad, crop-and-pad, random crop to shape), gamma / hue jitter, noise and
point removal are not implemented: configuring one raises NotImplementedError at construction (no silent fallback)."""
import ctypes

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
        self.normalized_image_range = normalized_image_range
        self.do_random_brightness = -1 not in random_brightness
        self.random_brightness = random_brightness
        self.do_random_contrast = -1 not in random_contrast
        self.random_contrast = random_contrast
        self.do_random_saturation = -1 not in random_saturation
        self.random_saturation = random_saturation
        unsupported = {
            'random_gamma': -1 not in random_gamma, 'random_hue': -1 not in random_hue,
            'random_noise': random_noise_type != 'none' and random_noise_spread > -1,
            'random_remove_patch_percent_range': -1 not in random_remove_patch_percent_range,
            'random_crop_to_shape': -1 not in random_crop_to_shape, 'random_rotate_max': random_rotate_max > 0,
            'random_crop_and_pad': -1 not in random_crop_and_pad, 'random_resize_and_crop': -1 not in random_resize_and_crop,
            'random_resize_and_pad': -1 not in random_resize_and_pad,
        }
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError('Transforms options without a native kernel: %s (DESIGN.md, scope table f2)' % ', '.join(bad))
        self.do_photometric_transforms = self.do_random_brightness or self.do_random_contrast or self.do_random_saturation
        self.do_image_normalization = normalized_image_range is not None
        self.do_random_horizontal_flip = 'horizontal' in random_flip_type
        self.do_random_vertical_flip = 'vertical' in random_flip_type
        self._norm = self._normalisation(normalized_image_range)
        self.rand_device = None        # where the draws are made; None = the images' device, as in the reference

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

    def transform(self, images_arr, intrinsics_arr=[], padding_modes=['constant'], interpolation_modes=['nearest'],
                  random_transform_probability=0.00):
        images_arr = list(images_arr)
        device = images_arr[0].device
        if device.type != 'cuda':
            raise RuntimeError('the augmentations run on CUDA only (no CPU fallback); got device %s' % device)
        if images_arr[0].ndim != 4:
            raise ValueError('Unsupported number of dimensions: {}'.format(images_arr[0].ndim))
        n_batch, n_channel = images_arr[0].shape[:2]
        do_random_transform = self._rand(n_batch, device) <= random_transform_probability                       # :229-230
        flags, factors = {}, {}
        for name, enabled, rng, ge in (('b', self.do_random_brightness, self.random_brightness, True),
                                       ('c', self.do_random_contrast, self.random_contrast, False),
                                       ('s', self.do_random_saturation, self.random_saturation, False)):
            if not enabled:
                continue
            roll = self._rand(n_batch, device)
            # brightness is applied when its roll is >= 0.5 (:243-245), the others when it is <= 0.5 (:256-258, :295-297)
            flags[name] = torch.logical_and(do_random_transform, roll >= 0.50 if ge else roll <= 0.50).to(torch.uint8).contiguous()
            values = self._rand(n_batch, device)
            lo, hi = rng
            factors[name] = ((hi - lo) * values + lo).contiguous()
        if self.do_photometric_transforms or self.do_image_normalization:
            images_arr = [self._photometric(im, flags, factors) for im in images_arr]
        else:
            images_arr = [im.float() for im in images_arr]
        if n_channel == 1:
            images_arr = [im[..., 0:1, :, :] for im in images_arr]
        do_h = do_v = None
        if self.do_random_horizontal_flip:                                                                       # :386-394
            do_h = torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8).contiguous()
        if self.do_random_vertical_flip:                                                                         # :396-404
            do_v = torch.logical_and(do_random_transform, self._rand(n_batch, device) <= 0.50).to(torch.uint8).contiguous()
        if do_h is not None or do_v is not None:
            images_arr = [self._flip(im, do_h, do_v) for im in images_arr]
        outputs = []
        if len(images_arr) > 0:
            outputs.append(images_arr)
        if len(intrinsics_arr) > 0:
            outputs.append(list(intrinsics_arr))
        return outputs[0] if len(outputs) == 1 else outputs

    def _photometric(self, images, flags, factors):
        if images.shape[1] != 3:
            if self.do_photometric_transforms:
                raise NotImplementedError('photometric transforms of %d-channel tensors' % images.shape[1])
        images = images.float().contiguous()
        n, c, h, w = images.shape
        if c != 3:                                  # normalisation only (single-channel input to a Transforms without jitter)
            mode, mean, std = self._norm
            if mode == 3:
                raise NotImplementedError('standard normalisation of %d-channel tensors' % c)
            return images / 255.0 if mode == 1 else (2.0 * (images / 255.0) - 1.0 if mode == 2 else images)
        out = torch.empty_like(images)
        mode, mean, std = self._norm
        ws = torch.empty(n, dtype=torch.int64, device=images.device) if 'c' in flags else None
        check(_lib.lib().ptta_augment_photometric(
            ptr(images), ptr(out), n, h, w, ptr(flags.get('b')), ptr(factors.get('b')), ptr(flags.get('c')), ptr(factors.get('c')),
            ptr(flags.get('s')), ptr(factors.get('s')), 1 if self.do_photometric_transforms else 0, mode,
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
