#!/usr/bin/env python
"""Parity report of the native NLSPN TTA step against the CPU oracle (oracle/nlspn_oracle.py): activations block by block,
losses, gradients of the 88 adapted tensors, adapted tensors after Adam.  Usage: nlspn_net_check.py [n h w [steps]]"""
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nlspn_oracle as NO           # checker
from oracle import msgchn_oracle as O
from tta_depth_completion_b200.nlspn_engine import NlspnEngine


def nrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nchw(t):
    return t.float().permute(0, 3, 1, 2)


def time_only(n, h, w, dev):
    """full-size timing of the eager step (no oracle): CUDA events around 5 steps after 2 warm-up steps, phases separately"""
    sd = NO.make_synthetic_checkpoint(0)
    eng = NlspnEngine(sd, n, h, w, dev)
    frames = []
    for t in range(3):
        image, sparse, _ = NO.synthetic_frame(5, t, n, h, w, 'kitti')
        frames.append((NO.normalize_image(image).to(dev), image.to(dev), sparse.to(dev)))
    for t in range(2):
        eng.tta_step(*frames[t % 3], 3e-4)
    torch.cuda.synchronize()
    print('losses', eng.read_losses(), 'mem GB %.2f' % (torch.cuda.max_memory_allocated() / 1e9))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    l0 = eng.launches
    ev[0].record()
    for t in range(5):
        eng.tta_step(*frames[t % 3], 3e-4)
    ev[1].record()
    torch.cuda.synchronize()
    print('ms/step %.3f  launches/step %d' % (ev[0].elapsed_time(ev[1]) / 5, (eng.launches - l0) // 5))
    # phases
    im_n, im_r, sp = frames[0]
    d_c, d_f, v_f = eng.B['clamped_depth'], eng.B['filtered_depth'], eng.B['filtered_validity']
    ev[0].record(); fe = eng.encoder('r.', im_n, d_c); eng.fe = fe
    ev[1].record(); eng.decoder(fe, d_c)
    ev[2].record(); fz = eng.encoder('z.', None, d_c)
    ev[3].record(); eng.emb = eng.mlp('z.', 'pred', eng.mlp('z.', 'proj', fz[-1].view(eng.R, 512))); eng.ref = eng.mlp('r.', 'proj_t', fe[-1].view(eng.R, 512))
    eng.loss(im_r, d_f, v_f, 80.0, 1.0, 1.0, 0.1)
    ev[4].record(); eng.backward()
    ev[5].record()
    torch.cuda.synchronize()
    for i, nm in enumerate(('encoder', 'decoder+prop', 'zero encoder', 'heads+loss', 'backward')):
        print('  %-14s %.3f ms' % (nm, ev[i].elapsed_time(ev[i + 1])))


def main():
    n, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1, 48, 80)
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    dev = torch.device('cuda:0')
    if os.environ.get('TIME_ONLY'):
        return time_only(n, h, w, dev)
    sd = NO.make_synthetic_checkpoint(0)
    sd_ref = {k: v.clone() for k, v in sd.items()}
    sd_emu0 = {k: v.clone() for k, v in sd.items()}
    eng = NlspnEngine(sd, n, h, w, dev)
    names = NO.adapt_parameter_names(sd_ref, 'meta_bn')
    assert names == eng.adapt_names
    state = O.AdamState(names, sd_ref)
    lr = 3e-4
    for t in range(steps):
        image, sparse, _ = NO.synthetic_frame(5, t, n, h, w, 'kitti')
        trace = {}
        ref = NO.tta_step(sd_ref, state, image, sparse, lr=lr, max_input_depth=80.0, return_grads=True, trace=trace)
        eng.tta_step(NO.normalize_image(image).to(dev), image.to(dev), sparse.to(dev), lr)
        torch.cuda.synchronize()
        got = eng.read_losses()
        print('--- step %d' % t)
        for k in ('loss', 'loss_sparse_depth', 'loss_smooth', 'loss_cos'):
            print('%-18s native %.6f oracle %.6f rel %.2e' % (k, got[k], ref[k], abs(got[k] - ref[k]) / max(abs(ref[k]), 1e-12)))
        B = eng.B
        pairs = [('fe1', nchw(B['r.fe1'])), ('fe2', nchw(B['r.conv2.2.bn2.act'])), ('fe3', nchw(B['r.conv3.3.bn2.act'])),
                 ('fe4', nchw(B['r.conv4.5.bn2.act'])), ('fe5', nchw(B['r.conv5.2.bn2.act'])), ('fe6', nchw(B['r.conv6.1.act'])),
                 ('fd5', nchw(B['r.dec5.1.act'])), ('fd4', nchw(B['r.dec4.1.act'])), ('fd3', nchw(B['r.dec3.1.act'])), ('fd2', nchw(B['r.dec2.1.act'])),
                 ('pred_init', B['pred_init']), ('guide', B['guide']), ('confidence', B['confidence']), ('offset', B['offset']), ('aff', B['aff']),
                 ('y', B['y']), ('fe6_zero', nchw(B['z.conv6.1.act']))]
        for k, v in pairs:
            print('  %-10s %.3e' % (k, nrel(v, trace[k].detach())))
        print('  %-10s %.3e' % ('emb', nrel(eng.emb.float(), trace['emb'].detach())))
        print('  %-10s %.3e' % ('ref', nrel(eng.ref.float(), trace['ref'].detach())))
        print('  %-10s %.3e' % ('output', nrel(B['output'], ref['output_depth'])))
        assert torch.equal(B['filtered_validity'].cpu(), ref['validity']) and torch.equal(B['filtered_depth'].cpu(), ref['sparse_depth'])
        if t == 0:
            sd_e = {k: v.clone() for k, v in sd_emu0.items()}
            emu = NO.tta_step(sd_e, O.AdamState(names, sd_e), image, sparse, lr=lr, max_input_depth=80.0, pr=O.Precision('bf16'), return_grads=True)
            ratio = []
            for k in names:
                gr = ref['grads'][k]
                e_nat = float((eng.grads[k].cpu() - gr).norm()) / max(float(gr.norm()), 1e-30)
                e_emu = float((emu['grads'][k] - gr).norm()) / max(float(gr.norm()), 1e-30)
                ratio.append((e_nat / max(e_emu, 1e-6), k, e_nat, e_emu))
            ratio.sort(reverse=True)
            print('native gradient error / error of the oracle with bf16 emulation: worst 8, median ratio %.2f' % ratio[len(ratio) // 2][0])
            for r, k, a, b in ratio[:8]:
                print('  %-34s native %.3e emulated %.3e ratio %.2f' % (k, a, b, r))
            worst = []
            gmax = max(float(g.norm()) for g in ref['grads'].values())
            for k in names:
                e = float((eng.grads[k].cpu() - ref['grads'][k]).norm())
                worst.append((e / max(float(ref['grads'][k].norm()), 1e-6 * gmax), k, float(ref['grads'][k].norm())))
            worst.sort(reverse=True)
            print('gradient errors (norm-wise rel), worst 12 of %d; median %.3e' % (len(worst), worst[len(worst) // 2][0]))
            for e, k, gn in worst[:12]:
                print('  %-34s %.3e (|g| %.3e)' % (k, e, gn))
    werr = sorted(((nrel(eng.params[k], sd_ref[k]), k) for k in names), reverse=True)
    print('adapted tensors after %d steps: worst %s median %.3e' % (steps, ['%s %.2e' % (k, e) for e, k in werr[:4]], werr[len(werr) // 2][0]))
    print('launches per step ~', eng.launches // steps)
    # timing of one step at this size
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(3):
        eng.tta_step(NO.normalize_image(image).to(dev), image.to(dev), sparse.to(dev), lr)
    torch.cuda.synchronize()
    print('ms/step (eager, incl. host): %.2f' % ((time.perf_counter() - t0) / 3 * 1e3))


if __name__ == '__main__':
    main()
