"""GPU: stage 2 of the source-domain preparation on the NLSPN back-end (tta_depth_completion_b200/nlspn_prepare.py: NlspnHeadTrainer, every
launch through the C ABI) against (a) the fixtures written by the REAL reference's stage-2 loop (oracle/gen_golden_nlspn_prepare.py;
src/head_main.py:259-278, 437-480), (b) the fp32 oracle and (c) the oracle's bf16 emulation (same rounding points as the native path).

Stated tolerances: losses within 5e-3 of the reference's (bf16 operands, fp32 accumulation); embedding rows and every gradient within 2.5 x the emulation's own
error + 3e-2 of the fp32 gradient, norm-wise; trained tensors within the distance the emulation itself ends up from the reference (x 2.5)
plus 5 % of the update; the heads' BatchNorm buffers within 2.5 x the emulation's error + 2e-2; the EMA copy within 2.5 x the emulation's distance + 1e-5."""
import glob
import os

import pytest
import torch

from oracle import msgchn_oracle as O
from oracle import nlspn_oracle as NO
from golden_util import GOLDEN_DIR, load_golden, rel, nrel

pytestmark = pytest.mark.gpu
NAMES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, 'nlspn_prep_head_*.pt')))
NOISE_GRAD = ('proj.0.bias', 'proj.3.bias', 'pred.0.bias')      # analytically zero gradients (a constant in front of a train-mode BatchNorm)
S = 128
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'nlspn_prepare_parity_report.txt')


def report(line):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, 'a') as f:
        f.write(line + '\n')


def initial_state(case):
    from tta_depth_completion_b200.nlspn_prepare import fresh_head_state
    sd = {k: v.clone() for k, v in NO.make_synthetic_checkpoint(case['ckpt_seed']).items()}
    torch.manual_seed(case['seed'])
    sd.update(fresh_head_state())
    return sd


def make_trainer(sd, n, h, w, dev):
    from tta_depth_completion_b200.nlspn_engine import NlspnEngine
    from tta_depth_completion_b200.nlspn_prepare import NlspnHeadTrainer
    eng = NlspnEngine(sd, n, h, w, dev, syncbn=False)
    return NlspnHeadTrainer(eng)


@pytest.mark.parametrize('name', NAMES)
def test_head_training_against_reference_fixture_and_oracle(name):
    fx = load_golden(name)
    case = fx['case']
    n, h, w = case['n'], case['h'], case['w']
    dev = torch.device('cuda:0')
    sd = initial_state(case)
    sd_ref = {k: v.clone() for k, v in sd.items()}
    sd_emu = {k: v.clone() for k, v in sd.items()}
    w0 = {k: sd[k].clone() for k in fx['trained']}
    tr = make_trainer(sd, n, h, w, dev)
    names = fx['trained']
    assert list(tr.params.keys()) == names
    st_ref, st_emu = O.AdamState(names, sd_ref), O.AdamState(names, sd_emu)
    for t, want in enumerate(fx['steps']):
        image, sparse, _ = NO.synthetic_frame(case['seq'], t, n, h, w, case['dataset'])
        tr.head_step(NO.normalize_image(image).to(dev), sparse.to(dev), case['lr'], max_input_depth=case['cap'])
        loss = tr.read_loss()
        ref = NO.head_step(sd_ref, st_ref, NO.normalize_image(image), torch.clamp(sparse, 0, case['cap']), lr=case['lr'], return_grads=True)
        emu = NO.head_step(sd_emu, st_emu, NO.normalize_image(image), torch.clamp(sparse, 0, case['cap']), lr=case['lr'], return_grads=True,
                           pr=O.Precision('bf16'))
        report('%s step %d loss native %.6f reference %.6f oracle %.6f emulation %.6f' % (name, t, loss, want['loss'], ref['loss'], emu['loss']))
        assert rel(loss, want['loss']) < 5e-3, (t, loss, want['loss'])
        for what, got, ex in (('emb', tr.emb, emu['emb']), ('ref', tr.ref, emu['ref'])):
            e_nat, e_emu = nrel(got[:4].float().cpu(), want[what + '_rows']), nrel(ex[:4], want[what + '_rows'])
            report('%s step %d %s rows native %.3e emulation %.3e' % (name, t, what, e_nat, e_emu))
            assert e_nat < 2.5 * e_emu + 5e-3, (t, what, e_nat, e_emu)
        if t > 0:
            continue                                             # later steps start from weights that already differ by rounding
        for k in names:
            if k in NOISE_GRAD:
                continue
            gr = ref['grads'][k]
            e_nat = float((tr.grads[k].cpu() - gr).norm()) / max(float(gr.norm()), 1e-30)
            e_emu = float((emu['grads'][k] - gr).norm()) / max(float(gr.norm()), 1e-30)
            report('%s grad %-16s native %.3e emulation %.3e (norm %.3e, reference %.3e)' % (name, k, e_nat, e_emu, float(gr.norm()), want['grad_norm'][k]))
            assert e_nat < 2.5 * e_emu + 3e-2, (k, e_nat, e_emu)
    torch.cuda.synchronize()
    for k in names:
        got = tr.params[k].cpu().flatten()[::S]
        want = fx['params_after_s128'][k]
        if k in NOISE_GRAD:
            assert float((got - want).abs().max()) <= 2.2 * case['lr'] * case['steps'], k
            continue
        upd = float((want - w0[k].flatten()[::S]).norm())
        e_nat = float((got - want).norm())
        e_emu = float((sd_emu[k].flatten()[::S] - want).norm())
        report('%s weights %-16s native %.3e emulation %.3e of the update' % (name, k, e_nat / max(upd, 1e-30), e_emu / max(upd, 1e-30)))
        assert e_nat <= 2.5 * e_emu + 0.05 * upd, (k, e_nat, e_emu, upd)
    for k, v in fx['buffers_after'].items():
        if k.endswith('num_batches_tracked'):
            assert int(tr.eng.sd[k]) == int(v), k
        else:
            e_nat, e_emu = nrel(tr.eng.sd[k].cpu(), v), nrel(sd_emu[k], v)
            report('%s buffer %-22s native %.3e emulation %.3e' % (name, k, e_nat, e_emu))
            assert e_nat < 2.5 * e_emu + 2e-2, (k, e_nat, e_emu)
    for k, v in fx['proj_t_after_s128'].items():
        e_nat, e_emu = nrel(tr.eng.sd[k].cpu().flatten()[::S], v), nrel(sd_emu[k].flatten()[::S], v)
        assert e_nat < 2.5 * e_emu + 1e-5, (k, e_nat, e_emu)          # EMA of a proj that is itself trained: (1 - tau) x its rounding distance


def test_head_training_at_frame_size():
    """1x352x1216 (1 672 rows of 512 features): two native steps against the fp32 oracle -- the loss of both steps, the gradient norms of
    the first"""
    dev = torch.device('cuda:0')
    case = dict(ckpt_seed=0, seed=99, n=1, h=352, w=1216, dataset='kitti', cap=80.0, lr=1e-3, seq=7)
    sd = initial_state(case)
    sd_ref = {k: v.clone() for k, v in sd.items()}
    tr = make_trainer(sd, 1, 352, 1216, dev)
    names = list(tr.params.keys())
    st_ref = O.AdamState(names, sd_ref)
    for t in range(2):
        image, sparse, _ = NO.synthetic_frame(case['seq'], t, 1, 352, 1216, 'kitti')
        tr.head_step(NO.normalize_image(image).to(dev), sparse.to(dev), case['lr'], max_input_depth=case['cap'])
        loss = tr.read_loss()
        ref = NO.head_step(sd_ref, st_ref, NO.normalize_image(image), torch.clamp(sparse, 0, case['cap']), lr=case['lr'], return_grads=True)
        report('fullsize step %d loss native %.6f oracle %.6f' % (t, loss, ref['loss']))
        assert rel(loss, ref['loss']) < 5e-3, (t, loss, ref['loss'])
        if t == 0:
            for k in names:
                if k in NOISE_GRAD:
                    continue
                gn, rn = float(tr.grads[k].norm()), float(ref['grads'][k].norm())
                e = float((tr.grads[k].cpu() - ref['grads'][k]).norm()) / max(rn, 1e-30)
                report('fullsize grad %-16s native norm %.4e oracle %.4e error %.3e' % (k, gn, rn, e))
                assert e < 8e-2, (k, e)


def test_head_training_through_the_wrapper():
    """the reference driver's stage-2 set-up lines (src/head_main.py:259-278) on the drop-in class, then the fused `head_step`: per-step losses
    against the reference's fixture, state_dict() holds the trained heads"""
    from tta_depth_completion_b200.external_model_adapt import ExternalModel_Adapt
    fx = load_golden(NAMES[-1])
    case = fx['case']
    n, h, w = case['n'], case['h'], case['w']
    dev = torch.device('cuda:0')
    m = ExternalModel_Adapt(model_name='nlspn', min_predict_depth=0.0, max_predict_depth=100.0, max_input_depth=case['cap'], offset=True,
                            dataset_name='kitti', device=dev)
    m._prepare_head(NO.PREPARE_MODE)
    m.load_state_dict(NO.make_synthetic_checkpoint(case['ckpt_seed']))
    torch.manual_seed(case['seed'])
    m.prepare_parameters('head_selfsup_ema')
    m.convert_syncbn()
    m.train(prepare=True)
    mean, std = torch.tensor(NO.IMAGENET_MEAN), torch.tensor(NO.IMAGENET_STD)
    m.set_image_normalization((1.0 / (255.0 * std)).tolist(), (-mean / std).tolist())
    sd0 = {k: v.detach().clone().cpu() for k, v in m.model.state_dict().items() if k.startswith(('proj.', 'pred.'))}
    for t, want in enumerate(fx['steps']):
        image, sparse, _ = NO.synthetic_frame(case['seq'], t, n, h, w, case['dataset'])
        m.head_step(image.to(dev), sparse.to(dev), case['lr'])
        loss = m.last_losses()['loss']
        assert rel(loss, want['loss']) < 5e-3, (t, loss, want['loss'])
    sd1 = m.model.state_dict()
    for k in fx['trained']:
        if k in NOISE_GRAD:
            continue
        got = sd1[k].detach().cpu().flatten()[::S]
        upd = float((fx['params_after_s128'][k] - sd0[k].flatten()[::S]).norm())
        assert float((got - fx['params_after_s128'][k]).norm()) < 0.6 * upd, k          # bf16 operands, few rows: see the report of the test above
        assert float((got - sd0[k].flatten()[::S]).norm()) > 0.5 * upd, k                # and it did move
    with pytest.raises(NotImplementedError):
        m.prepare_parameters('init_meta')


def test_captured_head_step_equals_eager():
    """graph=True (frame copied into trainer-owned staging buffers, step captured on the second call and replayed) against the eager
    launches, a fresh input tensor every step: same losses and the same trained tensors bit for bit"""
    dev = torch.device('cuda:0')
    case = dict(ckpt_seed=2, seed=77, n=1, h=64, w=128, dataset='kitti', cap=80.0, lr=1e-3, seq=9)
    trainers = [make_trainer(initial_state(case), 1, 64, 128, dev) for _ in range(2)]
    stream = torch.cuda.Stream()
    for t in range(5):
        image, sparse, _ = NO.synthetic_frame(case['seq'], t, 1, 64, 128, 'kitti')
        image, sparse = NO.normalize_image(image).to(dev), sparse.to(dev)
        trainers[0].head_step(image, sparse, case['lr'], max_input_depth=case['cap'])
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            trainers[1].head_step(image.clone(), sparse.clone(), case['lr'], max_input_depth=case['cap'], graph=True)
        torch.cuda.current_stream().wait_stream(stream)
        assert trainers[0].read_loss() == trainers[1].read_loss(), t
    assert trainers[1]._graph is not None
    for k in trainers[0].params:
        assert torch.equal(trainers[0].params[k], trainers[1].params[k]), k
        assert torch.equal(trainers[0].eng.sd['proj_t.0.weight'], trainers[1].eng.sd['proj_t.0.weight'])


def test_preparation_forwards_without_a_kernel_path_fail_loudly():
    from tta_depth_completion_b200.external_model_adapt import ExternalModel_Adapt
    dev = torch.device('cuda:0')
    m = ExternalModel_Adapt(model_name='nlspn', min_predict_depth=0.0, max_predict_depth=100.0, max_input_depth=80.0, offset=True,
                            dataset_name='kitti', device=dev)
    m._prepare_head(NO.PREPARE_MODE)
    m.load_state_dict(NO.make_synthetic_checkpoint(0))
    image, sparse, _ = NO.synthetic_frame(3, 0, 1, 32, 64, 'kitti')
    for lt in ('head_meta_selfsup_seq_ema_reverse', 'init_meta_selfsup_seq_ema'):
        with pytest.raises(NotImplementedError):
            m.forward(image=NO.normalize_image(image).to(dev), sparse_depth=sparse.to(dev), intrinsics=None, loss_type=lt)
    with pytest.raises(RuntimeError, match='prepare_parameters'):
        m.set_image_normalization((1.0,) * 3, (0.0,) * 3)
        m.head_step(image.to(dev), sparse.to(dev), 1e-3)
