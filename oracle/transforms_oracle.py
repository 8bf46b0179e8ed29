"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's `Transforms.transform` (src/transforms.py:192-668) for the options
the native path implements: brightness / contrast / saturation jitter (:236-301, per-sample torchvision tensor ops, :714-837), image
normalisation (:669-712) and per-sample flips (:386-407, 990-1034).  The third-party arithmetic is torchvision's
(`torchvision.transforms.functional.adjust_*`, pinned 0.10.1 by the reference's README.md:74; 0.26 here -- the tensor code path
`_blend` / `rgb_to_grayscale` is unchanged between the two) and is called as the reference calls it.  Pinned against fixtures produced by
the reference's own class: oracle/gen_golden_transforms.py, tests/test_transforms_oracle.py."""
import torch
from torchvision.transforms import functional


def draws(n_batch, cfg, probability, rand=lambda n: torch.rand(n)):
    """the reference's random draws, in its order (T:229-331, 386-404); cfg: dict with optional 'brightness' / 'contrast' /
    'saturation' ranges and 'flip' = subset of ('horizontal', 'vertical')"""
    d = {'do': rand(n_batch) <= probability}
    for name, ge in (('brightness', True), ('contrast', False), ('saturation', False)):
        if name in cfg:
            roll = rand(n_batch)
            d['do_' + name] = torch.logical_and(d['do'], roll >= 0.50 if ge else roll <= 0.50)
            lo, hi = cfg[name]
            d['f_' + name] = (hi - lo) * rand(n_batch) + lo
    for name in ('horizontal', 'vertical'):
        if name in cfg.get('flip', ()):
            d['do_' + name] = torch.logical_and(d['do'], rand(n_batch) <= 0.50)
    return d


def apply(images_arr, cfg, d, normalized_image_range=None):
    images_arr = [im.clone() for im in images_arr]
    photometric = any(k in cfg for k in ('brightness', 'contrast', 'saturation'))
    if photometric:
        images_arr = [im.to(torch.uint8) if torch.is_floating_point(im) else im for im in images_arr]          # T:236-240
    for name, fn in (('brightness', functional.adjust_brightness), ('contrast', functional.adjust_contrast),
                     ('saturation', functional.adjust_saturation)):
        if name in cfg:
            for images in images_arr:
                for b in range(images.shape[0]):
                    if d['do_' + name][b]:
                        images[b, ...] = fn(images[b], d['f_' + name][b])                                        # T:714-837
    images_arr = [im.float() for im in images_arr]
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
    for name, dim in (('horizontal', -1), ('vertical', -2)):
        if 'do_' + name in d:
            for images in images_arr:
                for b in range(images.shape[0]):
                    if d['do_' + name][b]:
                        images[b, ...] = torch.flip(images[b], dims=[dim])                                       # T:990-1034
    return images_arr
