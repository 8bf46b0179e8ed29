"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/nlspn_prep_head_*.pt by running the REAL reference's stage-2 loop (predictor heads,
src/head_main.py:259-275 construction, :437-480 step) on its NLSPN back-end (`ExternalModel_Adapt('nlspn')` from /root/reference through
oracle/ref_shims.py, network external_src/NLSPN/src/model/nlspnmodel_adapt.py):

    _prepare_head(prepare_mode) -> load the seeded checkpoint -> prepare_parameters('head_selfsup_ema') (re-creates proj / proj_t / pred and
    returns proj.* + pred.*) -> torch.optim.Adam -> convert_syncbn() -> per step: train(prepare=True), forward
    loss_type='head_meta_selfsup_seq_ema_reverse' (= _rgbd_meta_contrast_prepare: both encoders under no_grad with BatchNorm in eval mode,
    EMA copy of proj, emb = pred(proj(fe6 of the zero image)), ref = proj_t(fe6 of the frame)), compute_loss(loss_type='prepare'),
    backward, Adam step.

    python oracle/gen_golden_nlspn_prepare.py          # needs /root/reference; CPU, ~1 min

Shims: those of oracle/gen_golden_nlspn_net.py (none touches arithmetic).  The driver neither filters outliers nor augments here
(augmentation probability 0).  The freshly created heads are drawn from torch's global RNG: the fixture stores the seed, and the generator
asserts that the package's own constructor (nlspn_prepare.fresh_head_state) reproduces the reference's tensors bit for bit under that seed.
The fixture holds per-step losses and gradient norms, strided samples of the trained tensors / Adam moments / EMA copy after the last step,
and the heads' BatchNorm buffers."""
import contextlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims                                        # noqa: E402
from oracle import nlspn_oracle as NO                               # noqa: E402
from oracle import msgchn_oracle as O                               # noqa: E402
from oracle.gen_golden_nlspn_net import build_reference_nlspn       # noqa: E402
from tta_depth_completion_b200.nlspn_prepare import fresh_head_state     # noqa: E402  (constructor only: CPU tensors)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
CASES = [
    dict(name='nlspn_prep_head_kitti_2x64x96', ckpt_seed=0, n=2, h=64, w=96, dataset='kitti', cap=80.0, lr=1e-3, steps=3, seq=41, seed=5321),
    dict(name='nlspn_prep_head_kitti_1x48x160', ckpt_seed=1, n=1, h=48, w=160, dataset='kitti', cap=80.0, lr=1e-3, steps=2, seq=42, seed=5322),
]
HEAD_LOSS_TYPE = 'head_meta_selfsup_seq_ema_reverse'


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run_case(case):
    model, ref = build_reference_nlspn(case['cap'])                             # incl. _prepare_head(prepare_mode), head_main.py:259
    net = model.model.model
    sd_ckpt = NO.make_synthetic_checkpoint(case['ckpt_seed'])
    net.load_state_dict(sd_ckpt, strict=True)                                   # :266 restore_model minus torch.load
    torch.manual_seed(case['seed'])
    params = quiet(model.prepare_parameters, 'head_selfsup_ema')                # :268 (re-creates proj / proj_t / pred)
    named = {id(p): k for k, p in net.named_parameters()}
    handed = [named[id(p)] for p in params]
    assert tuple(handed) == NO.HEAD_TRAINED, handed
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    torch.manual_seed(case['seed'])
    mine = fresh_head_state()
    for k, v in mine.items():
        assert torch.equal(v, sd0[k]), 'constructor mismatch for %s' % k
    assert set(mine) == {k for k in sd0 if k.startswith(('proj', 'pred'))}
    opt = torch.optim.Adam(params, lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)     # :270-275
    ref_shims.enable_cpu_syncbn()
    model.convert_syncbn()                                                      # :278
    steps = []
    for t in range(case['steps']):
        model.train(prepare=True)
        image, sparse, dense = NO.synthetic_frame(case['seq'], t, case['n'], case['h'], case['w'], case['dataset'])
        out, emb, refm = model.forward(image=NO.normalize_image(image), sparse_depth=sparse, intrinsics=None, loss_type=HEAD_LOSS_TYPE)
        assert out is None
        loss, _ = model.compute_loss(input_rgb=image, output_depth=out, validity_map=None, ground_truth=dense, embedding=emb, reference=refm,
                                     loss_type='prepare')
        opt.zero_grad()
        loss.backward()
        pd = dict(net.named_parameters())
        steps.append({'loss': float(loss.detach()), 'grad_norm': {k: float(pd[k].grad.norm()) for k in handed},
                      'emb_rows': emb.detach()[:4].clone(), 'ref_rows': refm.detach()[:4].clone()})
        opt.step()
    sd1 = net.state_dict()
    st = opt.state_dict()['state']
    idx = {k: i for i, k in enumerate(handed)}
    encoder_buffers_untouched = all(torch.equal(sd1[k], sd0[k]) for k in sd0 if k.endswith(('running_mean', 'running_var'))
                                    and not k.startswith(('proj.', 'pred.')))
    assert encoder_buffers_untouched                                            # eval-mode BatchNorm everywhere but proj / pred
    return {'case': case, 'trained': handed, 'steps': steps, 'digest': O.checkpoint_digest(sd_ckpt),
            'digest_heads0': O.checkpoint_digest({k: sd0[k] for k in mine}),
            # every 128th element (the full tensors would be 50 MB per fixture)
            'params_after_s128': {k: sd1[k].flatten()[::128].clone() for k in handed},
            'update_norm': {k: float((sd1[k] - sd0[k]).norm()) for k in handed},
            'params_after_norm': {k: float(sd1[k].norm()) for k in handed},
            'exp_avg_s128': {k: st[idx[k]]['exp_avg'].flatten()[::128].clone() for k in handed},
            'exp_avg_sq_s128': {k: st[idx[k]]['exp_avg_sq'].flatten()[::128].clone() for k in handed},
            'proj_t_after_s128': {k: sd1[k].flatten()[::128].clone() for k in sd1 if k.startswith('proj_t.') and k.endswith(('weight', 'bias'))},
            'buffers_after': {k: sd1[k].clone() for k in sd1 if k.startswith(('proj', 'pred')) and k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))},
            'torch_version': torch.__version__}


def main():
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for case in CASES:
        if only and not any(o in case['name'] for o in only):
            continue
        fx = run_case(case)
        path = os.path.join(GOLDEN_DIR, case['name'] + '.pt')
        torch.save(fx, path)
        print('%-32s losses %s -> %s (%.0f KB)' % (case['name'], ' '.join('%.6f' % s['loss'] for s in fx['steps']),
                                                   os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
