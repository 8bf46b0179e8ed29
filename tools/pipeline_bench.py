"""The adaptation protocol from FILES to adapted weights on one B200 (SURVEY.md section 8 f2 + the step): how fast the whole per-frame
pipeline of src/tta_main.py runs when every stage is the library's own --

  PNG bytes (RGB 8-bit + 16-bit depth, KITTI-size 375 x 1242)
    -> host decode, `ops.decode_png_rgb8 / decode_png_gray16` on a thread pool (ctypes releases the GIL), straight into pinned memory
    -> H2D of the compact uint8 / uint16 frame on a copy stream (3.2 x fewer bytes than fp32)
    -> `ops.input_stage` (fp32 image / sparse depth / validity + the reference's bottom crop to 352 x 1216, src/datasets.py:83-170)
    -> `Transforms` (photometric jitter + normalisation; geometric set of the shipped scripts, bash/adapt/adapt_msgchn_vkitti.sh)  [--augment]
    -> `ExternalModel_Adapt.tta_step(graph=True)` (outlier removal, forward, losses, backward, Adam)

Reported: frames/s of (a) the decode pool alone, (b) the GPU side alone fed from already decoded pinned frames, (c) the whole pipeline,
and the same three for PIL as the decoder (what the reference's loaders use, src/data_utils.py:149-152,186).  Frames are synthetic
(tta_depth_completion_b200/synthetic.py) and encoded once with PIL at zlib level 6; the frame set (default 32 files) is cycled.

    python tools/pipeline_bench.py [--frames 32] [--steps 400] [--threads 0 (= all cores)] [--augment]"""
import argparse
import io
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tta_depth_completion_b200 import ExternalModel_Adapt, Transforms, ops, synthetic          # noqa: E402

H0, W0, H, W = 375, 1242, 352, 1216
MODE, CAP, LR = 'meta_selfsup_seq_2layers_ema', 80.0, 1e-4


def make_files(count):
    """(rgb_png_bytes, depth_png_bytes) per frame: KITTI-size files around the synthetic 352 x 1216 scene"""
    from PIL import Image
    files = []
    for t in range(count):
        image, sparse, _ = synthetic.synthetic_frame(1, t, 1, H0, W0, 'kitti')
        rgb = image[0].permute(1, 2, 0).clamp(0, 255).to(torch.uint8).numpy()
        d16 = (sparse[0, 0] * 256.0).clamp(0, 65535).to(torch.int32).numpy().astype(np.uint16)
        a, b = io.BytesIO(), io.BytesIO()
        Image.fromarray(rgb).save(a, format='PNG', compress_level=6)
        Image.fromarray(d16).save(b, format='PNG', compress_level=6)
        files.append((a.getvalue(), b.getvalue()))
    return files


def decode_native(f, rgb_out, d_out):
    ops.decode_png_rgb8(f[0], out=rgb_out)
    ops.decode_png_gray16(f[1], out=d_out)


def decode_pil(f, rgb_out, d_out):
    from PIL import Image
    rgb_out.numpy()[...] = np.asarray(Image.open(io.BytesIO(f[0])).convert('RGB'))
    d_out.numpy()[...] = np.array(Image.open(io.BytesIO(f[1])))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=32)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--augment', action='store_true')
    args = ap.parse_args()
    threads = args.threads or (os.cpu_count() or 1)
    dev = torch.device('cuda', 0)
    files = make_files(args.frames)
    out = {'frames': args.frames, 'steps': args.steps, 'decode_threads': threads, 'augment': bool(args.augment),
           'file_bytes_per_frame': int(np.mean([len(a) + len(b) for a, b in files])), 'h2d_bytes_per_frame': H0 * W0 * 3 + H0 * W0 * 2}

    # pinned ring of decoded frames: DEPTH slots, each one uint8 HWC image + one uint16 depth map
    DEPTH = 2 * threads + 4
    ring_rgb = [torch.empty((H0, W0, 3), dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]
    ring_d = [torch.empty((H0, W0), dtype=torch.uint16).pin_memory() for _ in range(DEPTH)]

    model = ExternalModel_Adapt('msg_chn', 0.0, 100.0, max_input_depth=CAP, device=dev)
    model._prepare_head(MODE)
    model.load_state_dict(synthetic.make_synthetic_checkpoint(0, MODE))
    model.set_image_normalization((1 / 255.0,) * 3, (0.0,) * 3)
    model.train()
    photometric = Transforms(normalized_image_range=[0, 255], random_brightness=[0.6, 1.4], random_contrast=[0.6, 1.4], random_saturation=[0.6, 1.4])
    geometric = Transforms(random_flip_type=['horizontal'], random_rotate_max=5, random_resize_and_crop=[1.0, 1.5])
    compute, copy = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage = [(torch.empty((1, H0, W0, 3), dtype=torch.uint8, device=dev), torch.empty((1, H0, W0), dtype=torch.uint16, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    copied = [None] * DEPTH              # per ring slot: event recorded when its H2D copies are done (the slot may then be decoded into again)

    def gpu_side(slot, k):
        """H2D of ring slot `slot` through device staging pair k, then input stage (+ augmentations) and the step"""
        with torch.cuda.stream(copy):
            copy.wait_event(consumed[k])
            stage[k][0][0].copy_(ring_rgb[slot], non_blocking=True)
            stage[k][1][0].copy_(ring_d[slot], non_blocking=True)
            ready[k].record(copy)
            ev = torch.cuda.Event()
            ev.record(copy)
            copied[slot] = ev
        with torch.cuda.stream(compute):
            compute.wait_event(ready[k])
            image, sparse, validity = ops.input_stage(stage[k][0], stage[k][1], crop_shape=(H, W), crop_type=('bottom',))
            consumed[k].record(compute)
            if args.augment:
                image, sparse = geometric.transform(images_arr=[image, sparse], interpolation_modes=['bilinear', 'nearest'], random_transform_probability=1.0)
                [image] = photometric.transform(images_arr=[image], random_transform_probability=1.0)
            model.tta_step(image, sparse, LR, 1.0, 1.0, 0.1, graph=True)

    def run(decoder, steps, with_gpu, with_decode):
        pool = ThreadPoolExecutor(max_workers=threads)
        for e in consumed:
            e.record(compute)
        for s_ in range(DEPTH):
            copied[s_] = None
        futs = {}

        def submit(i):
            slot = i % DEPTH
            if copied[slot] is not None:         # the frame that used this slot has left for the device
                copied[slot].synchronize()
                copied[slot] = None
            futs[i] = pool.submit(decoder, files[i % len(files)], ring_rgb[slot], ring_d[slot]) if with_decode else None
        ahead = min(DEPTH - 2, steps)            # decodes in flight ahead of the consumer
        for i in range(ahead):
            submit(i)
        nxt = ahead
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(steps):
            f = futs.pop(i)
            if f is not None:
                f.result()
            if with_gpu:
                gpu_side(i % DEPTH, i % 2)
            if nxt < steps:
                submit(nxt)
                nxt += 1
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        pool.shutdown()
        return steps / dt

    # warm-up: engine creation, graph capture, thread pool start
    run(decode_native, 8, True, True)
    for name, dec in (('native', decode_native), ('pil', decode_pil)):
        out['decode_only_%s_fps' % name] = round(run(dec, args.steps, False, True), 1)
    out['gpu_only_fps'] = round(run(decode_native, args.steps, True, False), 1)
    for name, dec in (('native', decode_native), ('pil', decode_pil)):
        out['pipeline_%s_fps' % name] = round(run(dec, args.steps, True, True), 1)
    out['losses'] = model.last_losses()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
