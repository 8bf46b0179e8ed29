"""Timing of the source-domain preparation steps on one B200: ptta_msgchn_init_step / ptta_msgchn_head_step at the
benchmark frame size, and the Linear weight-gradient GEMM (gemm_tn_tc_kernel + reduce) alone.  CUDA events, after warm-up.
    python tools/prepare_timing.py [--batch 1 4]"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tta_depth_completion_b200 import ExternalModel_Adapt, _lib                          # noqa: E402
from tta_depth_completion_b200.synthetic import synthetic_frame, get_checkpoint           # noqa: E402

DEV = 'cuda'


def timed(fn, iters, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, nargs='*', default=[1, 4])
    ap.add_argument('--iters', type=int, default=50)
    args = ap.parse_args()
    out = {}
    L = _lib.lib()
    for rows in (26752, 4 * 26752):
        a = torch.randn(rows, 512, device=DEV).bfloat16()
        b = torch.randn(rows, 512, device=DEV).bfloat16()
        c = torch.empty(512, 512, device=DEV)
        ws = torch.empty(L.ptta_gemm_tn_workspace_bytes(rows, 512, 512), dtype=torch.uint8, device=DEV)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        ms = timed(lambda: L.ptta_gemm_tn_bf16_tc(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(ws), rows, 512, 512, st), 200)
        ref = timed(lambda: torch.matmul(a.t(), b), 200)
        out['gemm_tn_%dx512x512' % rows] = {'us': round(ms * 1e3, 2), 'tflops': round(2.0 * rows * 512 * 512 / ms / 1e9, 1),
                                            'algorithmic_GBps': round((2 * rows * 512 * 2 + 512 * 512 * 4) / ms / 1e6, 1),
                                            'cublas_bf16_us': round(ref * 1e3, 2)}
    # on-device augmentations (csrc/augment.cuh): all three photometric transforms on, [0,1] normalisation; and one flip pass of the image
    for n in args.batch:
        h, w = 352, 1216
        img = (torch.rand(n, 3, h, w, device=DEV) * 255).contiguous()
        o = torch.empty_like(img)
        on = torch.ones(n, dtype=torch.uint8, device=DEV)
        f = torch.full((n,), 1.2, device=DEV)
        ws = torch.empty(n, dtype=torch.int64, device=DEV)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = _lib.ptr
        ms = timed(lambda: L.ptta_augment_photometric(P(img), P(o), n, h, w, P(on), P(f), P(on), P(f), P(on), P(f), None, None, None, None, None, None, 0.0, 0,
                                                          1, 1, None, None, P(ws), st), 200)
        alg = 3 * img.numel() * 4          # grey-sum pass reads the image, the apply pass reads it again and writes the result
        out['augment_photometric_%dx352x1216' % n] = {'us': round(ms * 1e3, 2), 'algorithmic_GBps': round(alg / ms / 1e6, 1)}
        ms = timed(lambda: L.ptta_augment_flip(P(img), P(o), n, 3, h, w, P(on), None, st), 200)
        out['augment_flip_%dx3x352x1216' % n] = {'us': round(ms * 1e3, 2), 'algorithmic_GBps': round(2 * img.numel() * 4 / ms / 1e6, 1)}
    for n in args.batch:
        sd = get_checkpoint('kitti_2layers_a', 'meta_selfsup_seq_2layers_ema')
        frames = [tuple(t.to(DEV) for t in synthetic_frame(60, k, n, 352, 1216, 'kitti')) for k in range(4)]
        for stage in ('init', 'head'):
            model = ExternalModel_Adapt('msg_chn', 0.0, 100.0, max_input_depth=80.0, device=torch.device(DEV))
            model._prepare_head('meta_selfsup_seq_2layers_ema')
            model.load_state_dict(sd)
            torch.manual_seed(1)
            if stage == 'head':
                model.prepare_parameters('head_selfsup_ema')
            model.set_image_normalization((1 / 255.0,) * 3, (0.0,) * 3)
            model.train()
            k = [0]

            def step():
                im, sp, gt = frames[k[0] % 4]
                k[0] += 1
                if stage == 'init':
                    model.init_step(im, sp, gt, 1e-3)
                else:
                    model.head_step(im, sp, 1e-3)
            ms = timed(step, args.iters)
            stream = torch.cuda.Stream()

            def gstep():
                im, sp, gt = frames[k[0] % 4]
                k[0] += 1
                if stage == 'init':
                    model.init_step(im, sp, gt, 1e-3, graph=True)
                else:
                    model.head_step(im, sp, 1e-3, graph=True)
            with torch.cuda.stream(stream):
                ms_graph = timed(gstep, args.iters)
            eng = model.model._engine_for(frames[0][0])
            l0 = eng.launch_count()
            step()
            out['%s_step_%dx352x1216' % (stage, n)] = {'ms': round(ms, 3), 'frames_per_s': round(n * 1e3 / ms, 1), 'ms_graph': round(ms_graph, 3), 'launches': eng.launch_count() - l0,
                                                       'loss': round(model.last_losses()['loss'], 5)}
    # stage 2 on the NLSPN back-end (nlspn_prepare.NlspnHeadTrainer: Python-orchestrated launches over the C ABI, eager)
    from tta_depth_completion_b200.synthetic import make_nlspn_checkpoint, IMAGENET_MEAN, IMAGENET_STD
    for n in args.batch:
        model = ExternalModel_Adapt('nlspn', 0.0, 100.0, max_input_depth=80.0, offset=True, device=torch.device(DEV))
        model._prepare_head('meta_selfsup_seq_1layer_ema')
        model.load_state_dict(make_nlspn_checkpoint(0))
        torch.manual_seed(1)
        model.prepare_parameters('head_selfsup_ema')
        model.set_image_normalization([1.0 / (255.0 * s_) for s_ in IMAGENET_STD], [-m_ / s_ for m_, s_ in zip(IMAGENET_MEAN, IMAGENET_STD)])
        frames = [tuple(t.to(DEV) for t in synthetic_frame(60, k, n, 352, 1216, 'kitti')) for k in range(4)]
        k = [0]

        def nstep():
            im, sp, gt = frames[k[0] % 4]
            k[0] += 1
            model.head_step(im, sp, 1e-3)
        ms = timed(nstep, args.iters)
        tr = model._last_engine
        l0 = tr.launches + tr.eng.launches
        nstep()
        n_launch = tr.launches + tr.eng.launches - l0
        stream = torch.cuda.Stream()

        def ngstep():
            im, sp, gt = frames[k[0] % 4]
            k[0] += 1
            model.head_step(im, sp, 1e-3, graph=True)
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            ms_graph = timed(ngstep, args.iters)
        torch.cuda.current_stream().wait_stream(stream)
        out['nlspn_head_step_%dx352x1216' % n] = {'ms': round(ms, 3), 'frames_per_s': round(n * 1e3 / ms, 1), 'ms_graph': round(ms_graph, 3),
                                                   'frames_per_s_graph': round(n * 1e3 / ms_graph, 1), 'launches': n_launch,
                                                   'loss': round(model.last_losses()['loss'], 5)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
