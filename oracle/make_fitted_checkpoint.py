"""TEST INFRASTRUCTURE ONLY -- produces tests/golden/ckpt_*.pt: MSG-CHN ProxyTTA checkpoints FITTED on the synthetic scene by
running the REAL reference (/root/reference through oracle/ref_shims.py), so that the parity fixtures are not taken in the
degenerate regime of a randomly initialised network (output ~ 0, loss_cos ~ 2, MAE ~ 40 m).

    python oracle/make_fitted_checkpoint.py            # CPU, ~10 min; then re-run oracle/gen_golden.py

The three stages are the reference's own source-domain preparation in miniature (no checkpoint ships with the reference:
README.md:144-149 points at Google Drive):
  0. base network, supervised: stands in for the upstream MSG-CHN training the reference downloads.  The reference model's
     `forward(loss_type='pretrain')` / `compute_loss(loss_type='pretrain')` (src/msg_chn_model_adapt.py:224-264: masked L2
     against the dense ground truth), torch.optim.Adam over `model.parameters()`.
  1. stage 1, meta-layer initialisation (src/init_main.py:288-311, 448-572): `prepare_parameters(prepare_mode)` creates the
     meta layer and freezes everything else; forward `loss_type='init_meta_seq_ema'`, the same supervised loss.
  2. stage 2, proxy heads (src/head_main.py:259-275, 464-480): `_prepare_head(prepare_mode)`,
     `prepare_parameters('head_selfsup_ema')`, forward `loss_type='head_meta_selfsup_seq_ema_reverse'`
     (network_exp_msg_chn_adapt.py:609-699: pred(proj(zero-image latents)) against proj(real latents), with the EMA copy
     `proj_t` updated every step), loss `'prepare'` = cosine distance (src/external_model_adapt.py:524-540).
Two snapshots of stage 2 are kept: `_a` after a few head steps (loss_cos stays above the 0.3 gate of
src/external_model_adapt.py:424 at test time) and `_b` after the full head fit (loss_cos < 0.3: the gate fires, w_cos = 0).
`_b` stores only the tensors that differ from `_a`.

Frames: crops of the oracle's synthetic sequences (oracle/msgchn_oracle.py: synthetic_frame), sequence seeds >= 100 (the
fixtures and the benchmark use seeds < 100).
"""
import argparse
import contextlib
import io
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import msgchn_oracle as O          # noqa: E402
from oracle import ref_shims                   # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

CONFIGS = {
    # name: prepare_mode, dataset, depth cap, full frame size the crops are cut from
    'kitti_2layers': dict(prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti', cap=80.0, full=(352, 1216)),
    'void_1layer': dict(prepare_mode='meta_selfsup_seq_1layer_ema', dataset='void', cap=8.0, full=(480, 640)),
}


def batch_of_crops(cfg, step, n, ch, cw, density=None):
    """n crops (ch x cw) of full-size synthetic frames; half of them native small frames (the small fixtures use those)."""
    g = torch.Generator().manual_seed(9000 + step)
    H, W = cfg['full']
    ims, sps, dns = [], [], []
    for i in range(n):
        seq = 100 + int(torch.randint(0, 50, (1,), generator=g))
        t = int(torch.randint(0, 200, (1,), generator=g))
        if i % 2 == 0:
            image, sparse, dense = O.synthetic_frame(seq, t, 1, H, W, cfg['dataset'])
            y0 = int(torch.randint(0, H - ch + 1, (1,), generator=g))
            x0 = int(torch.randint(0, W - cw + 1, (1,), generator=g))
            image, sparse, dense = (a[:, :, y0:y0 + ch, x0:x0 + cw] for a in (image, sparse, dense))
        else:
            image, sparse, dense = O.synthetic_frame(seq, t, 1, ch, cw, cfg['dataset'])
        if density is not None:      # VOID's 0.5 % leaves ~40 points in a small crop: train with a denser sample as well
            m = (torch.rand(dense.shape, generator=g) < density).float()
            sparse = dense * m
        ims.append(image); sps.append(sparse); dns.append(dense)
    return torch.cat(ims).contiguous(), torch.cat(sps).contiguous(), torch.cat(dns).contiguous()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def fit(name, base_steps, init_steps, head_steps_a, head_steps_b, batch, crop, log=print):
    cfg = CONFIGS[name]
    torch.manual_seed(1234)
    ref = ref_shims.load_reference()
    cwd = os.getcwd()
    os.chdir(ref_shims.REFERENCE_ROOT)
    try:
        model = quiet(ref.ExternalModel_Adapt, model_name='msg_chn', max_input_depth=cfg['cap'], min_predict_depth=0.0,
                      max_predict_depth=100.0, device=torch.device('cpu'), from_scratch=False, dataset_name='', offset=True)
    finally:
        os.chdir(cwd)
    outlier = ref.OutlierRemoval(7, 1.5)
    ch, cw = crop
    dens = 0.03 if cfg['dataset'] == 'void' else None

    def frames(step):
        image, sparse, dense = batch_of_crops(cfg, step, batch, ch, cw, dens if step % 2 else None)
        v = torch.where(sparse > 0, torch.ones_like(sparse), sparse)
        fsd, _ = outlier.remove_outliers(sparse_depth=sparse, validity_map=v)
        vgt = torch.where(dense > 0, torch.ones_like(dense), dense)
        return image, fsd, dense, vgt

    def supervised(params, steps, loss_type, lr, tag, step0):
        opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
        model.train(meta=True)
        t0 = time.time()
        for s in range(steps):
            if s == int(0.7 * steps):
                for gr in opt.param_groups:
                    gr['lr'] = lr * 0.3
            image, fsd, dense, vgt = frames(step0 + s)
            out = model.forward(image=image / 255.0, sparse_depth=fsd, intrinsics=None, loss_type=loss_type)
            loss, _ = model.compute_loss(input_rgb=image, output_depth=out, validity_map=vgt, ground_truth=dense,
                                         embedding=None, reference=None, dataset_name='', loss_type='pretrain')
            opt.zero_grad()
            loss.backward()
            opt.step()
            if s % 50 == 0 or s == steps - 1:
                with torch.no_grad():
                    mae = float((out[0] - dense).abs().mean())
                log('%s %-5s step %4d  L2 %.4f  MAE %.3f m  (%.0f s)' % (name, tag, s, float(loss), mae, time.time() - t0))

    # ---- stage 0: base network ----
    supervised(model.parameters(), base_steps, 'pretrain', 2e-3, 'base', 0)
    # ---- stage 1: meta layer (init_main.py) ----
    init_mode = cfg['prepare_mode'].replace('_selfsup', '').replace('_ema', '')        # 'meta_seq_2layers': no heads yet
    metaparams = quiet(model.prepare_parameters, init_mode)
    supervised(metaparams, init_steps, 'init_meta_seq_ema', 1e-3, 'init', 100000)
    # ---- stage 2: proxy heads (head_main.py) ----
    net = model.model.model
    meta_state = {k: v.clone() for k, v in net.state_dict().items() if 'meta' in k}
    quiet(model._prepare_head, cfg['prepare_mode'])           # head_main.py:259 (re-creates the meta layer: restored next, as :267 does)
    net.load_state_dict(meta_state, strict=False)
    for p in net.parameters():
        p.requires_grad_(True)
    head_params = quiet(model.prepare_parameters, 'head_selfsup_ema')      # head_main.py:268
    opt = torch.optim.Adam(head_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    snaps = {}
    t0 = time.time()

    def held_out_loss_cos():
        """loss_cos the TTA step would see on fixture-like frames (sequence seeds < 100), train-mode BatchNorm as in adaptation"""
        keep = {k: v.clone() for k, v in net.state_dict().items()}
        tot = 0.0
        with torch.no_grad():
            for q in range(3):
                image, sparse, dense = O.synthetic_frame(1 + q, q, 1, ch, cw, cfg['dataset'])
                v = torch.where(sparse > 0, torch.ones_like(sparse), sparse)
                fsd, _ = outlier.remove_outliers(sparse_depth=sparse, validity_map=v)
                _, emb, refm = model.forward(image=image / 255.0, sparse_depth=fsd, intrinsics=None,
                                             loss_type='adapt_meta_selfsup_seq_ema_reverse')
                tot += float(model.compute_loss(input_rgb=image, embedding=emb, reference=refm, loss_type='prepare')[0])
        net.load_state_dict(keep)
        return tot / 3

    for s in range(head_steps_b):
        model.train(prepare=True)
        image, fsd, dense, vgt = frames(200000 + s)
        out, emb, refm = model.forward(image=image / 255.0, sparse_depth=fsd, intrinsics=None,
                                       loss_type='head_meta_selfsup_seq_ema_reverse')
        loss, _ = model.compute_loss(input_rgb=image, output_depth=out, validity_map=vgt, ground_truth=dense, embedding=emb,
                                     reference=refm, loss_type='prepare')
        opt.zero_grad()
        loss.backward()
        opt.step()
        if s % 25 == 0 or s == head_steps_b - 1:
            log('%s head  step %4d  loss_cos %.4f  (%.0f s)' % (name, s, float(loss), time.time() - t0))
        if s + 1 <= head_steps_a:        # 'a' = the last early snapshot whose held-out loss_cos is still well above the 0.3 gate
            ho = held_out_loss_cos()
            log('%s head  step %4d  held-out loss_cos %.4f' % (name, s, ho))
            if ho > 0.45 or 'a' not in snaps:
                snaps['a'] = {k: v.detach().clone() for k, v in net.state_dict().items()}
                snaps['a_info'] = (s + 1, ho)
    snaps['b'] = {k: v.detach().clone() for k, v in net.state_dict().items()}
    snaps['b_info'] = (head_steps_b, held_out_loss_cos())
    log('%s snapshot a: %d head steps, held-out loss_cos %.4f; b: %d steps, %.4f' % ((name,) + snaps['a_info'] + snaps['b_info']))
    assert snaps['a_info'][1] > 0.35 and snaps['b_info'][1] < 0.27, 'snapshots do not straddle the loss_cos < 0.3 gate'
    snaps = {k: v for k, v in snaps.items() if not k.endswith('_info')}
    want = O.make_synthetic_checkpoint(0, cfg['prepare_mode'])
    for tag, sd in snaps.items():
        assert list(sd.keys()) != [] and set(sd.keys()) == set(want.keys()), (set(sd) ^ set(want))
        for k in want:
            assert tuple(sd[k].shape) == tuple(want[k].shape), k
    return snaps


def save(name, snaps):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    a, b = snaps['a'], snaps['b']
    pa = os.path.join(GOLDEN_DIR, 'ckpt_%s_a.pt' % name)
    torch.save({'net': a, 'made_by': 'oracle/make_fitted_checkpoint.py', 'torch_version': torch.__version__}, pa)
    diff = {k: v for k, v in b.items() if not torch.equal(v, a[k])}
    pb = os.path.join(GOLDEN_DIR, 'ckpt_%s_b.pt' % name)
    torch.save({'net_delta': diff, 'base': os.path.basename(pa), 'made_by': 'oracle/make_fitted_checkpoint.py'}, pb)
    for p in (pa, pb):
        print('%s  %.1f MB' % (os.path.relpath(p, ROOT), os.path.getsize(p) / 1e6))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', nargs='*', default=list(CONFIGS))
    ap.add_argument('--base-steps', type=int, default=1500)
    ap.add_argument('--init-steps', type=int, default=200)
    ap.add_argument('--head-steps-a', type=int, default=30)
    ap.add_argument('--head-steps-b', type=int, default=400)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--crop', type=int, nargs=2, default=(64, 128))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    for name in args.configs:
        snaps = fit(name, args.base_steps, args.init_steps, args.head_steps_a, args.head_steps_b, args.batch, tuple(args.crop))
        save(name, snaps)


if __name__ == '__main__':
    main()
