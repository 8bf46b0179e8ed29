"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the REAL reference
(/root/reference, imported through oracle/ref_shims.py) on seeded synthetic checkpoints and frames.

    python oracle/gen_golden.py            # rewrites every fixture (needs /root/reference; CPU, ~1 min)

The driver lines executed per step are the reference's own (src/tta_main.py:583-590, 610-633):
validity map -> OutlierRemoval(7,1.5) -> model.forward(normalised image, filtered sparse) ->
model.compute_loss(raw image, ..., loss_type='adapt') -> zero_grad / backward / Adam.step, then an
eval-mode forward (src/tta_main.py:729-736).  Geometric / photometric augmentation is disabled
(probability 0), as in the measured configuration.

Each fixture stores: the case description (enough to regenerate checkpoint + frames from seeds),
a digest of the checkpoint, per-step losses, per-step L2 norms of every adapted tensor and its
gradient, the adapted tensors / Adam moments / BN buffers after the last step, the filtered sparse
depth + validity of the last step (bit-exact targets), the train-mode output depth of the last step,
samples of emb/ref and the eval-mode output after the last step.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import msgchn_oracle as O          # noqa: E402
from oracle import ref_shims                   # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

CASES = [
    # name, prepare_mode, dataset, N, H, W, steps, lr, max_input_depth, ckpt seed, sequence seed
    dict(name='msgchn_2layers_kitti_1x64x128', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti',
         n=1, h=64, w=128, steps=3, lr=1e-4, max_input_depth=80.0, ckpt_seed=0, seq_seed=1, density=None),
    dict(name='msgchn_2layers_kitti_2x48x80', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti',
         n=2, h=48, w=80, steps=2, lr=1e-4, max_input_depth=80.0, ckpt_seed=3, seq_seed=2, density=None),
    dict(name='msgchn_1layer_void_1x48x64', prepare_mode='meta_selfsup_seq_1layer_ema', dataset='void',
         n=1, h=48, w=64, steps=3, lr=3e-3, max_input_depth=8.0, ckpt_seed=1, seq_seed=5, density=0.03),
    # H, W not multiples of 16 -> exercises the pad + flip-pad ensembling of src/msg_chn_model_adapt.py:58-125
    dict(name='msgchn_2layers_kitti_1x40x72_pad', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti',
         n=1, h=40, w=72, steps=1, lr=1e-4, max_input_depth=80.0, ckpt_seed=0, seq_seed=7, density=None),
    # FITTED checkpoints (oracle/make_fitted_checkpoint.py: the reference's own supervised / head stages in miniature): the network
    # predicts depth (MAE ~0.5 m instead of ~40 m), loss_cos is O(0.1-1) instead of 2.0
    dict(name='msgchn_fit_kitti_1x64x128', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti', ckpt='kitti_2layers_a',
         n=1, h=64, w=128, steps=3, lr=1e-4, max_input_depth=80.0, ckpt_seed=None, seq_seed=21, density=None),
    # ... and with the heads fully fitted: loss_cos < 0.3, so the gate of src/external_model_adapt.py:424 sets w_cos = 0
    dict(name='msgchn_fit_kitti_gate_2x48x80', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti', ckpt='kitti_2layers_b',
         n=2, h=48, w=80, steps=3, lr=1e-4, max_input_depth=80.0, ckpt_seed=None, seq_seed=22, density=None),
    dict(name='msgchn_fit_void_1x48x64', prepare_mode='meta_selfsup_seq_1layer_ema', dataset='void', ckpt='void_1layer_a',
         n=1, h=48, w=64, steps=3, lr=3e-3, max_input_depth=8.0, ckpt_seed=None, seq_seed=23, density=0.03),
    # test-domain shift (what TTA is for): the sensor reads 3 % deeper than the surface the network was fitted on, so the L1 residual
    # (0.15 ... 2.3 m) is far above the bf16 noise of the prediction and sign(pred - d) is stable between implementations
    dict(name='msgchn_fit_kitti_shift_1x64x128', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti', ckpt='kitti_2layers_a',
         n=1, h=64, w=128, steps=3, lr=1e-4, max_input_depth=80.0, ckpt_seed=None, seq_seed=25, density=None, depth_scale=1.03),
    dict(name='msgchn_fit_void_shift_1x48x64', prepare_mode='meta_selfsup_seq_1layer_ema', dataset='void', ckpt='void_1layer_a',
         n=1, h=48, w=64, steps=3, lr=3e-3, max_input_depth=8.0, ckpt_seed=None, seq_seed=26, density=0.03, depth_scale=1.05),
    dict(name='msgchn_fit_kitti_1x40x72_pad', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti', ckpt='kitti_2layers_a',
         n=1, h=40, w=72, steps=2, lr=1e-4, max_input_depth=80.0, ckpt_seed=None, seq_seed=24, density=None),
]
W_SD, W_SM, W_COS = 1.0, 1.0, 0.1


def case_frame(case, t):
    image, sparse, dense = O.synthetic_frame(case['seq_seed'], t, case['n'], case['h'], case['w'], case['dataset'],
                                             depth_scale=case.get('depth_scale', 1.0))
    if case.get('density'):
        # tiny frames at 0.5 % density would have ~15 points; re-sample denser for the small fixtures
        g = torch.Generator().manual_seed(77 + t)
        mask = (torch.rand(dense.shape, generator=g) < case['density']).float()
        sparse = dense * mask
    return image, sparse, dense


def run_reference_case(case):
    ref = ref_shims.load_reference()
    sd0 = O.get_checkpoint(case.get('ckpt', case['ckpt_seed']), case['prepare_mode'])
    model = ref_shims.build_reference_msgchn(case['prepare_mode'], case['max_input_depth'])
    net = model.model.model
    missing = net.load_state_dict(sd0, strict=True)        # same keys/shapes as the reference checkpoint
    params = model.adapt_parameters(mode='meta')
    names = [k for k, p in net.named_parameters() if any(p is q for q in params)]
    optimizer = torch.optim.Adam(params=params, lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    outlier_removal = ref.OutlierRemoval(7, 1.5)
    loss_type = 'adapt_meta_selfsup_seq_ema_reverse'
    steps = []
    for t in range(case['steps']):
        image, sparse_depth, dense = case_frame(case, t)
        model.train()
        validity_map_depth = torch.where(sparse_depth > 0, torch.ones_like(sparse_depth), sparse_depth)
        fsd, fvm = outlier_removal.remove_outliers(sparse_depth=sparse_depth, validity_map=validity_map_depth)
        image1 = image / 255.0                            # normalized_image_range 0 1
        output_depth, emb, refm = model.forward(image=image1, sparse_depth=fsd, intrinsics=None, crop_mask=None,
                                                loss_type=loss_type)
        loss, info = model.compute_loss(input_rgb=image.detach(), output_depth=output_depth,
                                        sparse_depth=fsd.detach(), validity_map=fvm.detach(),
                                        embedding=emb, reference=refm,
                                        w_loss_sparse_depth=W_SD, w_loss_smoothness=W_SM, w_loss_cos=W_COS,
                                        loss_type='adapt')
        optimizer.zero_grad()
        loss.backward()
        grad_norms = {k: float(dict(net.named_parameters())[k].grad.norm()) for k in names}
        optimizer.step()
        steps.append({
            'loss': float(loss), 'loss_smooth': float(info['loss_smooth']),
            'loss_sparse_depth': float(info['loss_sparse_depth']), 'loss_cos': float(info['loss_cos']),
            'grad_norm': grad_norms,
            'param_norm': {k: float(dict(net.named_parameters())[k].detach().norm()) for k in names},
            'n_valid': int(fvm.sum()), 'n_valid_in': int(validity_map_depth.sum()),
            'w_cos_eff': 0.0 if float(info['loss_cos']) < 0.3 else W_COS,
        })
    model.eval()
    with torch.no_grad():
        eval_out = model.forward(image=image1, sparse_depth=fsd, intrinsics=None, crop_mask=None, loss_type=loss_type)
    sd_after = net.state_dict()
    opt_state = optimizer.state_dict()['state']
    fixture = {
        'case': case, 'digest': O.checkpoint_digest(sd0), 'adapt_names': names, 'steps': steps,
        'params_after': {k: sd_after[k].clone() for k in names},
        'exp_avg': {k: opt_state[i]['exp_avg'].clone() for i, k in enumerate(names)},
        'exp_avg_sq': {k: opt_state[i]['exp_avg_sq'].clone() for i, k in enumerate(names)},
        'buffers_after': {k: v.clone() for k, v in sd_after.items()
                          if k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))},
        'sparse_depth_filtered': fsd.clone(), 'validity_filtered': fvm.to(torch.uint8),
        'output_depth': output_depth.detach().clone(), 'eval_output_depth': eval_out.clone(),
        'emb_rows': emb.detach()[:4].clone(), 'ref_rows': refm.detach()[:4].clone(),
        'torch_version': torch.__version__,
    }
    return fixture


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]
    for case in CASES:
        if only and not any(o in case['name'] for o in only):
            continue
        fx = run_reference_case(case)
        path = os.path.join(GOLDEN_DIR, case['name'] + '.pt')
        torch.save(fx, path)
        last = fx['steps'][-1]
        print('%-36s steps=%d loss=%.6f sd=%.6f sm=%.6f cos=%.6f n_valid=%d/%d  -> %s (%.0f KB)' % (
            case['name'], len(fx['steps']), last['loss'], last['loss_sparse_depth'], last['loss_smooth'],
            last['loss_cos'], last['n_valid'], last['n_valid_in'], os.path.relpath(path, ROOT),
            os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
