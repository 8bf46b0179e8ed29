"""TEST INFRASTRUCTURE ONLY -- a checkpoint file WRITTEN BY THE REAL REFERENCE in the middle of a TTA run
(`MsgChnModel_Adapt.save_model`, /root/reference/src/msg_chn_model_adapt.py:518-545: {'net', 'optimizer', 'train_step'}), plus what the
reference does next from it (one more step: losses, adapted tensors, Adam moments).  tests/test_checkpoint_gpu.py restores the file
into the native classes (`restore_model`, incl. the torch.optim.Adam state and `train_step`) and continues the run.

    python oracle/gen_golden_ckpt.py        # needs /root/reference; CPU, seconds

The network weights are the seeded synthetic checkpoint (ckpt_seed 0), so the file is ~7 MB of fp32 that cannot be regenerated
on the GPU box; it is committed as written."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import msgchn_oracle as O          # noqa: E402
from oracle import ref_shims                   # noqa: E402
from oracle.gen_golden import case_frame, W_SD, W_SM, W_COS          # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
CASE = dict(name='refckpt_2layers_kitti_1x64x128', prepare_mode='meta_selfsup_seq_2layers_ema', dataset='kitti',
            n=1, h=64, w=128, steps_before=2, lr=1e-4, max_input_depth=80.0, ckpt_seed=0, seq_seed=31, density=None)


def main():
    torch.set_num_threads(os.cpu_count())
    case = CASE
    ref = ref_shims.load_reference()
    sd0 = O.get_checkpoint(case['ckpt_seed'], case['prepare_mode'])
    model = ref_shims.build_reference_msgchn(case['prepare_mode'], case['max_input_depth'])
    net = model.model.model
    net.load_state_dict(sd0, strict=True)
    params = model.adapt_parameters(mode='meta')
    names = [k for k, p in net.named_parameters() if any(p is q for q in params)]
    optimizer = torch.optim.Adam(params=params, lr=case['lr'], betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    outlier_removal = ref.OutlierRemoval(7, 1.5)

    def step(t):
        image, sparse_depth, _ = case_frame(case, t)
        model.train()
        validity = torch.where(sparse_depth > 0, torch.ones_like(sparse_depth), sparse_depth)
        fsd, fvm = outlier_removal.remove_outliers(sparse_depth=sparse_depth, validity_map=validity)
        out, emb, refm = model.forward(image=image / 255.0, sparse_depth=fsd, intrinsics=None, crop_mask=None,
                                       loss_type='adapt_meta_selfsup_seq_ema_reverse')
        loss, info = model.compute_loss(input_rgb=image.detach(), output_depth=out, sparse_depth=fsd.detach(), validity_map=fvm.detach(),
                                        embedding=emb, reference=refm, w_loss_sparse_depth=W_SD, w_loss_smoothness=W_SM,
                                        w_loss_cos=W_COS, loss_type='adapt')
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return {'loss': float(loss), 'loss_smooth': float(info['loss_smooth']), 'loss_sparse_depth': float(info['loss_sparse_depth']),
                'loss_cos': float(info['loss_cos'])}

    before = [step(t) for t in range(case['steps_before'])]
    ckpt_path = os.path.join(GOLDEN_DIR, case['name'] + '.ckpt.pth')
    model.save_model(ckpt_path, case['steps_before'], optimizer)          # the reference's own writer
    after = step(case['steps_before'])
    sd_after = net.state_dict()
    opt_state = optimizer.state_dict()['state']
    fx = {'case': case, 'adapt_names': names, 'steps_before': before, 'step_after': after,
          'params_after': {k: sd_after[k].clone() for k in names},
          'exp_avg_after': {k: opt_state[i]['exp_avg'].clone() for i, k in enumerate(names)},
          'exp_avg_sq_after': {k: opt_state[i]['exp_avg_sq'].clone() for i, k in enumerate(names)},
          'adam_step_after': int(opt_state[0]['step']),
          'buffers_after': {k: v.clone() for k, v in sd_after.items() if k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))},
          'torch_version': torch.__version__}
    path = os.path.join(GOLDEN_DIR, case['name'] + '.pt')
    torch.save(fx, path)
    ck = torch.load(ckpt_path, weights_only=False)
    print('%s: %d net entries, optimizer state for %d tensors (step %d), train_step %d, %.1f MB; next step loss %.6f' % (
        os.path.relpath(ckpt_path, ROOT), len(ck['net']), len(ck['optimizer']['state']), int(ck['optimizer']['state'][0]['step']),
        ck['train_step'], os.path.getsize(ckpt_path) / 1e6, after['loss']))


if __name__ == '__main__':
    main()
