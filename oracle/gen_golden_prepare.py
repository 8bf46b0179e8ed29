"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/prep_*.pt by running the REAL reference's source-domain preparation loops
(/root/reference through oracle/ref_shims.py) for a few steps on the fitted miniature checkpoint:

  stage 1, meta-layer initialisation (src/init_main.py:288-311, 482-522): `prepare_parameters('meta_seq_<k>')` creates the meta layer and
      freezes the rest, forward `loss_type='init_meta_seq_ema'`, `compute_loss(loss_type='pretrain')` (masked L2 against the dense
      ground truth, src/msg_chn_model_adapt.py:224-264), torch.optim.Adam over the meta tensors;
  stage 2, predictor head (src/head_main.py:259-275, 437-480): `_prepare_head(mode)`, restore, `prepare_parameters('head_selfsup_ema')`
      (re-creates proj / proj_t / pred), forward `loss_type='head_meta_selfsup_seq_ema_reverse'`, `compute_loss(loss_type='prepare')`
      (cosine distance, src/external_model_adapt.py:524-540), torch.optim.Adam over proj.* and pred.*.

    python oracle/gen_golden_prepare.py          # needs /root/reference; CPU, ~1 min

Neither driver filters outliers or augments here (augmentation probability 0).  The freshly created layers are drawn from torch's
global RNG: the fixture stores the seed and a digest, and the generator asserts that the package's own constructors
(external_model_adapt.add_head_state) reproduce the reference's tensors bit for bit under that seed, so nothing but the seed
has to be committed.  Each fixture holds per-step losses and gradient norms, the trained tensors / Adam moments / BatchNorm buffers / EMA
copy after the last step.
"""
import contextlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import msgchn_oracle as O          # noqa: E402
from oracle import ref_shims                   # noqa: E402
from tta_depth_completion_b200.external_model_adapt import add_head_state      # noqa: E402  (constructors only: CPU tensors)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

CASES = [
    dict(name='prep_init_kitti_2layers_2x48x80', stage='init', prepare_mode='meta_selfsup_seq_2layers_ema', init_mode='meta_seq_2layers',
         ckpt='kitti_2layers_a', dataset='kitti', n=2, h=48, w=80, steps=3, lr=1e-3, max_input_depth=80.0, seq_seed=31, seed=4321),
    dict(name='prep_init_void_1layer_1x48x64', stage='init', prepare_mode='meta_selfsup_seq_1layer_ema', init_mode='meta_seq_1layer',
         ckpt='void_1layer_a', dataset='void', n=1, h=48, w=64, steps=3, lr=1e-3, max_input_depth=8.0, seq_seed=32, seed=4322, density=0.03),
    dict(name='prep_head_kitti_2layers_2x48x80', stage='head', prepare_mode='meta_selfsup_seq_2layers_ema',
         ckpt='kitti_2layers_a', dataset='kitti', n=2, h=48, w=80, steps=3, lr=1e-3, max_input_depth=80.0, seq_seed=33, seed=4323),
    dict(name='prep_head_void_1layer_1x64x64', stage='head', prepare_mode='meta_selfsup_seq_1layer_ema',
         ckpt='void_1layer_a', dataset='void', n=1, h=64, w=64, steps=3, lr=1e-3, max_input_depth=8.0, seq_seed=34, seed=4324, density=0.03),
]
HEAD_LOSS_TYPE = 'head_meta_selfsup_seq_ema_reverse'
INIT_LOSS_TYPE = 'init_meta_seq_ema'


def case_frame(case, t):
    image, sparse, dense = O.synthetic_frame(case['seq_seed'], t, case['n'], case['h'], case['w'], case['dataset'])
    if case.get('density'):
        g = torch.Generator().manual_seed(77 + t)
        sparse = dense * (torch.rand(dense.shape, generator=g) < case['density']).float()
    return image, sparse, dense


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def fresh_reference(case):
    ref = ref_shims.load_reference()
    cwd = os.getcwd()
    os.chdir(ref_shims.REFERENCE_ROOT)
    try:
        model = quiet(ref.ExternalModel_Adapt, model_name='msg_chn', max_input_depth=case['max_input_depth'], min_predict_depth=0.0,
                      max_predict_depth=100.0, device=torch.device('cpu'), from_scratch=False, dataset_name='', offset=True)
    finally:
        os.chdir(cwd)
    return model


def buffers(sd):
    return {k: v.clone() for k, v in sd.items() if k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))}


def run_init_case(case):
    sd_full = O.get_checkpoint(case['ckpt'], case['prepare_mode'])
    base = {k: v for k, v in sd_full.items() if not k.startswith(('conv1_rgb_meta', 'proj', 'pred'))}
    model = fresh_reference(case)
    net = model.model.model
    net.load_state_dict(base, strict=True)
    torch.manual_seed(case['seed'])
    params = quiet(model.prepare_parameters, case['init_mode'])                  # init_main.py:288 / :304
    names = [k for k, p in net.named_parameters() if any(p is q for q in params)]
    # the package's constructor under the same seed == the reference's fresh meta layer
    torch.manual_seed(case['seed'])
    mine = add_head_state({}, case['init_mode'])
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    for k, v in mine.items():
        assert torch.equal(v, sd0[k]), 'constructor mismatch for %s' % k
    assert set(mine) == {k for k in sd0 if 'meta' in k}
    opt = torch.optim.Adam(params, lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    model.train(meta=True)
    steps = []
    for t in range(case['steps']):
        image, sparse, dense = case_frame(case, t)
        vgt = torch.where(dense > 0, torch.ones_like(dense), dense)
        out = model.forward(image=image / 255.0, sparse_depth=sparse, intrinsics=None, loss_type=INIT_LOSS_TYPE)
        loss, _ = model.compute_loss(input_rgb=image, output_depth=out, validity_map=vgt, ground_truth=dense, embedding=None, reference=None,
                                     dataset_name='', loss_type='pretrain')
        opt.zero_grad()
        loss.backward()
        pd = dict(net.named_parameters())
        steps.append({'loss': float(loss.detach()), 'grad_norm': {k: float(pd[k].grad.norm()) for k in names}})
        opt.step()
    sd1 = net.state_dict()
    st = opt.state_dict()['state']
    return {'case': case, 'trained': names, 'digest0': O.checkpoint_digest({k: sd0[k] for k in mine}), 'steps': steps,
            'params_after': {k: sd1[k].clone() for k in names},
            'exp_avg': {k: st[i]['exp_avg'].clone() for i, k in enumerate(names)},
            'exp_avg_sq': {k: st[i]['exp_avg_sq'].clone() for i, k in enumerate(names)},
            'buffers_after': buffers(sd1), 'output_depth': out[0].detach().clone(), 'torch_version': torch.__version__}


def run_head_case(case):
    sd_full = O.get_checkpoint(case['ckpt'], case['prepare_mode'])
    model = fresh_reference(case)
    net = model.model.model
    quiet(model._prepare_head, case['prepare_mode'])                            # head_main.py:259
    net.load_state_dict(sd_full, strict=True)                                   # :266 restore_model
    torch.manual_seed(case['seed'])
    params = quiet(model.prepare_parameters, 'head_selfsup_ema')                # :268 (re-creates proj / proj_t / pred)
    handed = [k for k, p in net.named_parameters() if any(p is q for q in params)]
    torch.manual_seed(case['seed'])
    mine = {}
    add_head_state(mine, 'head_selfsup_ema')
    add_head_state(mine, 'head_selfsup_ema')                                    # the reference builds the heads twice (W:295-298)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    for k, v in mine.items():
        assert torch.equal(v, sd0[k]), 'constructor mismatch for %s' % k
    opt = torch.optim.Adam(params, lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    steps = []
    for t in range(case['steps']):
        model.train(prepare=True)
        image, sparse, dense = case_frame(case, t)
        vgt = torch.where(dense > 0, torch.ones_like(dense), dense)
        out, emb, refm = model.forward(image=image / 255.0, sparse_depth=sparse, intrinsics=None, loss_type=HEAD_LOSS_TYPE)
        loss, _ = model.compute_loss(input_rgb=image, output_depth=out, validity_map=vgt, ground_truth=dense, embedding=emb, reference=refm,
                                     loss_type='prepare')
        opt.zero_grad()
        loss.backward()
        pd = dict(net.named_parameters())
        steps.append({'loss': float(loss.detach()), 'grad_norm': {k: float(pd[k].grad.norm()) for k in handed if pd[k].grad is not None}})
        opt.step()
    trained = list(steps[-1]['grad_norm'].keys())
    assert tuple(trained) == O.HEAD_TRAINED, trained                            # proj's output is detached (N:692): only pred is trained
    sd1 = net.state_dict()
    st = opt.state_dict()['state']
    idx = {k: i for i, k in enumerate(handed)}
    for k in handed:
        if k not in trained:
            assert idx[k] not in st and torch.equal(sd1[k], sd0[k]), k          # Adam never touched proj
    return {'case': case, 'handed_to_adam': handed, 'trained': trained, 'steps': steps,
            'digest0': O.checkpoint_digest({k: sd0[k] for k in mine}),
            'params_after': {k: sd1[k].clone() for k in trained},
            # every 16th element of the moments / every 8th of the EMA copy (the full tensors would be 5 MB per fixture)
            'exp_avg_s16': {k: st[idx[k]]['exp_avg'].flatten()[::16].clone() for k in trained},
            'exp_avg_sq_s16': {k: st[idx[k]]['exp_avg_sq'].flatten()[::16].clone() for k in trained},
            'proj_t_after_s8': {k: sd1[k].flatten()[::8].clone() for k in sd1 if k.startswith('proj_t.') and k.endswith(('weight', 'bias'))},
            'buffers_after': buffers(sd1), 'emb_rows': emb.detach()[:4].clone(), 'ref_rows': refm.detach()[:4].clone(),
            'torch_version': torch.__version__}


def main():
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for case in CASES:
        if only and not any(o in case['name'] for o in only):
            continue
        fx = run_init_case(case) if case['stage'] == 'init' else run_head_case(case)
        path = os.path.join(GOLDEN_DIR, case['name'] + '.pt')
        torch.save(fx, path)
        print('%-36s losses %s -> %s (%.0f KB)' % (case['name'], ' '.join('%.6f' % s['loss'] for s in fx['steps']),
                                                   os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
