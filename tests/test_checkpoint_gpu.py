"""Checkpoint round trip: a checkpoint file WRITTEN BY THE REAL REFERENCE in the middle of a TTA run
(tests/golden/refckpt_*.ckpt.pth, oracle/gen_golden_ckpt.py: `save_model` of src/msg_chn_model_adapt.py:518-545 -> {'net',
'optimizer', 'train_step'}) is restored into the native classes and the run is continued -- through the reference driver's own
calls with torch.optim.Adam AND through the fused `tta_step` (whose Adam moments / step counter are loaded from the same file) --
and must reproduce what the reference did next.  Then the native `save_model` output is read back: same keys, dtypes, shapes and
optimizer-state layout as the reference's file, bit-identical state after a second restore."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from golden_util import GOLDEN_DIR, load_golden, case_frame, rel, nrel, W_SD, W_SM, W_COS
from test_msgchn_step_gpu import make_model, ZERO_GRAD, weight_tolerance, TOL_LOSS
from oracle import msgchn_oracle as O

DEV = 'cuda'
NAME = 'refckpt_2layers_kitti_1x64x128'
CKPT = os.path.join(GOLDEN_DIR, NAME + '.ckpt.pth')


def _fresh_model(case):
    # a DIFFERENT random checkpoint (seed 9): everything the continued run needs must come from the restored file
    return make_model(case['prepare_mode'], O.make_synthetic_checkpoint(9, case['prepare_mode']), case['max_input_depth'])


def _driver_step(model, opt, image, sparse):
    from tta_depth_completion_b200 import OutlierRemoval
    model.train()
    validity = torch.where(sparse > 0, torch.ones_like(sparse), sparse)
    fsd, fvm = OutlierRemoval(7, 1.5).remove_outliers(sparse_depth=sparse, validity_map=validity)
    out, emb, ref = model.forward(image=image / 255.0, sparse_depth=fsd, intrinsics=None, crop_mask=None,
                                  loss_type='adapt_meta_selfsup_seq_ema_reverse')
    loss, info = model.compute_loss(input_rgb=image.detach(), output_depth=out, sparse_depth=fsd.detach(), validity_map=fvm.detach(),
                                    embedding=emb, reference=ref, w_loss_sparse_depth=W_SD, w_loss_smoothness=W_SM, w_loss_cos=W_COS,
                                    loss_type='adapt')
    opt.zero_grad()
    loss.backward()
    opt.step()
    return {'loss': float(loss), 'loss_smooth': float(info['loss_smooth']), 'loss_sparse_depth': float(info['loss_sparse_depth']),
            'loss_cos': float(info['loss_cos'])}


def _check_continuation(fx, got_losses, sd_after, m_after, v_after, step_after):
    for k in ('loss', 'loss_sparse_depth', 'loss_smooth', 'loss_cos'):
        assert rel(got_losses[k], fx['step_after'][k]) < TOL_LOSS, (k, got_losses[k], fx['step_after'][k])
    assert step_after == fx['adam_step_after']
    ref_ck = torch.load(CKPT, map_location='cpu', weights_only=False)
    for k in fx['adapt_names']:
        if k in ZERO_GRAD:
            continue
        upd = nrel(ref_ck['net'][k], fx['params_after'][k])                   # this ONE step's update (it starts from warm moments)
        e = nrel(sd_after[k].cpu(), fx['params_after'][k])
        assert e < weight_tolerance(upd), (k, e, upd)
        # moments after the third step: 2/3 restored from the file + 1/3 this step's gradient (bf16-path gradient error 2-7 % here)
        assert nrel(m_after[k].cpu(), fx['exp_avg_after'][k]) < 0.1, (k, nrel(m_after[k].cpu(), fx['exp_avg_after'][k]))
        assert nrel(v_after[k].cpu(), fx['exp_avg_sq_after'][k]) < 0.1, (k, nrel(v_after[k].cpu(), fx['exp_avg_sq_after'][k]))
    for k, v in fx['buffers_after'].items():
        if k.endswith('num_batches_tracked'):
            assert int(sd_after[k]) == int(v), k


def test_restore_reference_checkpoint_and_continue_with_torch_adam():
    fx = load_golden(NAME)
    case = fx['case']
    model = _fresh_model(case)
    opt = torch.optim.Adam(model.adapt_parameters('meta'), lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    opt, train_step = model.restore_model(CKPT, opt)
    assert train_step == case['steps_before']
    ref_ck = torch.load(CKPT, map_location='cpu', weights_only=False)
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref_ck['net'].keys())                      # same keys in the same order as the reference module
    for k, v in ref_ck['net'].items():
        assert torch.equal(sd[k].cpu(), v), k                                 # restored bit-exactly (incl. BN buffers, num_batches_tracked)
    for i, k in enumerate(fx['adapt_names']):
        st = opt.state[opt.param_groups[0]['params'][i]]
        assert torch.equal(st['exp_avg'].cpu(), ref_ck['optimizer']['state'][i]['exp_avg']), k
        assert torch.equal(st['exp_avg_sq'].cpu(), ref_ck['optimizer']['state'][i]['exp_avg_sq']), k
        assert int(st['step']) == case['steps_before']
    image, sparse, _ = case_frame(case, case['steps_before'])
    got = _driver_step(model, opt, image.to(DEV), sparse.to(DEV))
    params = opt.param_groups[0]['params']
    m = {k: opt.state[params[i]]['exp_avg'] for i, k in enumerate(fx['adapt_names'])}
    v = {k: opt.state[params[i]]['exp_avg_sq'] for i, k in enumerate(fx['adapt_names'])}
    _check_continuation(fx, got, model.state_dict(), m, v, int(opt.state[params[0]]['step']))


def test_restore_reference_checkpoint_and_continue_with_fused_step():
    """no torch optimiser at all: restore_model loads the file's Adam moments and step count into the fused Adam"""
    fx = load_golden(NAME)
    case = fx['case']
    model = _fresh_model(case)
    _, train_step = model.restore_model(CKPT)
    assert train_step == case['steps_before']
    assert model.model.adam_step_count() == case['steps_before']
    image, sparse, _ = case_frame(case, case['steps_before'])
    model.tta_step(image.to(DEV), sparse.to(DEV), case['lr'], W_SD, W_SM, W_COS)
    _check_continuation(fx, model.last_losses(), model.state_dict(), model.model._m_views, model.model._v_views, model.model.adam_step_count())


@pytest.mark.parametrize('fused', [False, True])
def test_save_model_writes_the_reference_format_and_round_trips(tmp_path, fused):
    fx = load_golden(NAME)
    case = fx['case']
    a = _fresh_model(case)
    opt = None if fused else torch.optim.Adam(a.adapt_parameters('meta'), lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    opt, _ = a.restore_model(CKPT, opt)
    image, sparse, _ = case_frame(case, case['steps_before'])
    image, sparse = image.to(DEV), sparse.to(DEV)
    if fused:
        a.tta_step(image, sparse, case['lr'], W_SD, W_SM, W_COS)
    else:
        _driver_step(a, opt, image, sparse)
    path = str(tmp_path / 'native.pth')
    a.save_model(path, case['steps_before'] + 1, opt)
    mine = torch.load(path, map_location='cpu', weights_only=False)
    ref_ck = torch.load(CKPT, map_location='cpu', weights_only=False)
    assert set(mine.keys()) == set(ref_ck.keys()) == {'net', 'optimizer', 'train_step'}
    assert mine['train_step'] == case['steps_before'] + 1
    assert list(mine['net'].keys()) == list(ref_ck['net'].keys())
    for k, v in ref_ck['net'].items():
        assert mine['net'][k].dtype == v.dtype and tuple(mine['net'][k].shape) == tuple(v.shape), k
    assert set(mine['optimizer'].keys()) == set(ref_ck['optimizer'].keys())
    assert set(mine['optimizer']['param_groups'][0].keys()) == set(ref_ck['optimizer']['param_groups'][0].keys())
    assert mine['optimizer']['param_groups'][0]['params'] == ref_ck['optimizer']['param_groups'][0]['params']
    for i in ref_ck['optimizer']['state']:
        assert set(mine['optimizer']['state'][i].keys()) == set(ref_ck['optimizer']['state'][i].keys())
        assert int(mine['optimizer']['state'][i]['step']) == case['steps_before'] + 1
    # second restore (other wrapper, torch optimiser): bit-identical weights, buffers, moments
    b = _fresh_model(case)
    opt_b = torch.optim.Adam(b.adapt_parameters('meta'), lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    opt_b, step_b = b.restore_model(path, opt_b)
    assert step_b == case['steps_before'] + 1
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    for i, k in enumerate(fx['adapt_names']):
        assert torch.equal(b.model._m_views[k].cpu(), mine['optimizer']['state'][i]['exp_avg']), k
        assert torch.equal(opt_b.state[opt_b.param_groups[0]['params'][i]]['exp_avg_sq'].cpu(), mine['optimizer']['state'][i]['exp_avg_sq']), k
    assert b.model.adam_step_count() == case['steps_before'] + 1
    # ... and the two wrappers take the same next step (fused on both: same kernels, same state -> bit-identical)
    image2, sparse2, _ = case_frame(case, case['steps_before'] + 1)
    if fused:
        a.tta_step(image2.to(DEV), sparse2.to(DEV), case['lr'], W_SD, W_SM, W_COS)
        b.tta_step(image2.to(DEV), sparse2.to(DEV), case['lr'], W_SD, W_SM, W_COS)
        for k in fx['adapt_names']:
            assert torch.equal(a.state_dict()[k], b.state_dict()[k]), k
