"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/nlspn_net_*.pt by running the REAL reference NLSPN back-end
(`ExternalModel_Adapt('nlspn')` from /root/reference/src, network external_src/NLSPN/src/model/nlspnmodel_adapt.py) through
the driver's own lines (src/tta_main.py:309-346 construction, :583-633 step) on a seeded checkpoint and seeded frames.

    python oracle/gen_golden_nlspn_net.py       # needs /root/reference; CPU, about a minute

Shims (none touches arithmetic): the ones of oracle/ref_shims.py, a stub for `skimage.restoration.inpaint`
(src/data_utils.py:24, eval-time only) and `modulated_deform_conv_func` resolved to oracle.nlspn_prop_oracle.MDConvFn (the
reference's DCN CUDA extension has no CPU path) and SyncBatchNorm.forward on CPU (ref_shims.enable_cpu_syncbn: world size 1 = F.batch_norm).
convert_syncbn() IS called, as the driver does (src/tta_main.py:327): it decides which tensors adapt_parameters('meta_bn') returns.  DDP is skipped.

The checkpoint is NOT stored (26 M parameters): both sides rebuild it from `nlspn_oracle.make_synthetic_checkpoint(seed)`;
the fixture keeps its digest, the key/shape manifest of the reference's state dict (oracle/nlspn_state_manifest.json), the
losses / output depth / embeddings of every step, the gradients of the adapted tensors at step 1 and the adapted tensors
after the last step."""
import contextlib
import io
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims                     # noqa: E402
from oracle import nlspn_oracle as NO            # noqa: E402
from oracle import nlspn_prop_oracle as P        # noqa: E402
from oracle import msgchn_oracle as O            # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
CASES = [
    dict(name='nlspn_net_kitti_1x48x80', seed=0, n=1, h=48, w=80, dataset='kitti', cap=80.0, lr=3e-4, steps=2, seq=5),
    dict(name='nlspn_net_kitti_2x32x64', seed=1, n=2, h=32, w=64, dataset='kitti', cap=80.0, lr=3e-4, steps=1, seq=6),
]


def build_reference_nlspn(max_input_depth):
    ref = ref_shims.load_reference()
    for name in ('skimage', 'skimage.restoration', 'skimage.restoration.inpaint'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['skimage.restoration.inpaint'].inpaint_biharmonic = None
    sys.modules['skimage.restoration'].inpaint = sys.modules['skimage.restoration.inpaint']
    sys.modules['skimage'].restoration = sys.modules['skimage.restoration']
    shim = types.ModuleType('modulated_deform_conv_func')
    shim.ModulatedDeformConvFunction = P.MDConvFn
    sys.modules['modulated_deform_conv_func'] = shim
    cwd = os.getcwd()
    os.chdir(ref_shims.REFERENCE_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.ExternalModel_Adapt(model_name='nlspn', max_input_depth=max_input_depth, min_predict_depth=0.0,
                                            max_predict_depth=100.0, device=torch.device('cpu'), from_scratch=False,
                                            dataset_name='', offset=True)
            model._prepare_head(NO.PREPARE_MODE)
    finally:
        os.chdir(cwd)
    return model, ref


def run_case(case):
    model, ref = build_reference_nlspn(case['cap'])
    net = model.model.model
    sd = NO.make_synthetic_checkpoint(case['seed'])
    manifest = {k: list(v.shape) for k, v in net.state_dict().items()}
    missing = [k for k in manifest if k not in sd or list(sd[k].shape) != manifest[k]]
    extra = [k for k in sd if k not in manifest]
    assert not missing and not extra, (missing, extra)
    net.load_state_dict(sd, strict=True)                                   # restore_model (W:418-440) minus torch.load
    ref_shims.enable_cpu_syncbn()
    model.convert_syncbn()                                                 # T:327 -- BatchNorm2d AND the heads' BatchNorm1d become SyncBatchNorm
    params = model.adapt_parameters('meta_bn')                             # T:339
    named = {id(p): k for k, p in net.named_parameters()}
    names = [named[id(p)] for p in params]
    assert names == NO.adapt_parameter_names(sd, 'meta_bn'), 'adapted-parameter order differs from the restatement'
    optimizer = torch.optim.Adam([{'params': params, 'weight_decay': 0.0}], lr=case['lr'], betas=(0.9, 0.999), eps=1e-8)
    outlier = ref.OutlierRemoval(7, 1.5)
    model.train()
    steps = []
    grads1 = None
    for t in range(case['steps']):
        image, sparse, dense = NO.synthetic_frame(case['seq'], t, case['n'], case['h'], case['w'], case['dataset'])
        validity = torch.where(sparse > 0, torch.ones_like(sparse), sparse)                       # T:583-586
        f_sparse, f_valid = outlier.remove_outliers(sparse_depth=sparse, validity_map=validity)  # T:589-590
        out, emb, refm = model.forward(image=NO.normalize_image(image), sparse_depth=f_sparse, intrinsics=None,
                                       loss_type='adapt_meta_selfsup_seq_ema_reverse')
        loss, info = model.compute_loss(input_rgb=image, output_depth=out, sparse_depth=f_sparse, validity_map=f_valid,
                                        embedding=emb, reference=refm, w_loss_sparse_depth=1.0, w_loss_smoothness=1.0,
                                        w_loss_cos=0.1, loss_type='adapt')
        optimizer.zero_grad()
        loss.backward()
        if t == 0:
            grads1 = {k: p.grad.detach().clone() for k, p in zip(names, params)}
        optimizer.step()
        steps.append(dict(loss=float(loss), loss_smooth=float(info['loss_smooth']), loss_sparse_depth=float(info['loss_sparse_depth']),
                          loss_cos=float(info['loss_cos']), output_depth=out.detach().clone(), emb=emb.detach().clone().half(),
                          ref=refm.detach().clone().half(), validity=f_valid.clone(), sparse_depth=f_sparse.clone()))
    adapted = {k: p.detach().clone() for k, p in zip(names, params)}
    return dict(case=case, digest=O.checkpoint_digest(sd), names=names, steps=steps, grads_step1=grads1, adapted=adapted), manifest


def main():
    manifest = None
    for case in CASES:
        fx, manifest = run_case(case)
        path = os.path.join(GOLDEN_DIR, case['name'] + '.pt')
        torch.save(fx, path)
        s = fx['steps'][-1]
        print('wrote %s (%.0f KB): loss %.5f sd %.5f sm %.5f cos %.5f, %d adapted tensors / %d elements' % (
            path, os.path.getsize(path) / 1e3, s['loss'], s['loss_sparse_depth'], s['loss_smooth'], s['loss_cos'], len(fx['names']),
            sum(v.numel() for v in fx['adapted'].values())))
    with open(os.path.join(HERE, 'nlspn_state_manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=0)


if __name__ == '__main__':
    main()
