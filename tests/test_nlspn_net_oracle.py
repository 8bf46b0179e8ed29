"""The CPU oracle of the NLSPN back-end (oracle/nlspn_oracle.py) against fixtures produced by
running the real reference (oracle/gen_golden_nlspn_net.py: `ExternalModel_Adapt('nlspn')`, adapt mode `meta_bn`, the driver's
forward / compute_loss / backward / Adam lines).  This is what pins that oracle: the reference has no golden vectors."""
import glob
import json
import os

import pytest
import torch

from oracle import msgchn_oracle as O
from oracle import nlspn_oracle as NO
from golden_util import GOLDEN_DIR, load_golden, rel, nrel

# fp32 CPU vs fp32 CPU of the same torch build: summation-order noise only
TOL_LOSS = 5e-5
TOL_TENSOR = 1e-4
NAMES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, 'nlspn_net_*.pt')))


def test_checkpoint_keys_match_reference_manifest():
    """key set and shapes of the seeded checkpoint == the reference's NLSPNModel_Adapt.state_dict() (manifest written by the
    generator from the live reference module)"""
    with open(os.path.join(os.path.dirname(GOLDEN_DIR), '..', 'oracle', 'nlspn_state_manifest.json')) as f:
        manifest = json.load(f)
    sd = NO.make_synthetic_checkpoint(0)
    assert set(sd) == set(manifest)
    for k, shape in manifest.items():
        assert list(sd[k].shape) == shape, k
    names = NO.adapt_parameter_names(sd, 'meta_bn')
    # 88 tensors / 40 048 elements from the meta conv + BatchNorm2d layers + the three heads' BatchNorm1d affine pairs, which
    # convert_syncbn() (src/tta_main.py:327) turns into SyncBatchNorm instances that adapt_parameters('meta_bn') matches as well
    assert len(names) == 94 and sum(sd[k].numel() for k in names) == 40048 + 6 * 1024
    assert names[-6:] == ['proj.1.weight', 'proj.1.bias', 'proj_t.1.weight', 'proj_t.1.bias', 'pred.1.weight', 'pred.1.bias']


@pytest.mark.parametrize('name', NAMES)
def test_oracle_matches_reference_fixture(name):
    fx = load_golden(name)
    case = fx['case']
    sd = NO.make_synthetic_checkpoint(case['seed'])
    assert O.checkpoint_digest(sd) == pytest.approx(fx['digest'], rel=1e-12), 'seeded checkpoint differs'
    names = NO.adapt_parameter_names(sd, 'meta_bn')
    assert names == fx['names']
    state = O.AdamState(names, sd)
    for t in range(case['steps']):
        image, sparse, _ = NO.synthetic_frame(case['seq'], t, case['n'], case['h'], case['w'], case['dataset'])
        res = NO.tta_step(sd, state, image, sparse, lr=case['lr'], max_input_depth=case['cap'], return_grads=True)
        g = fx['steps'][t]
        for k in ('loss', 'loss_smooth', 'loss_sparse_depth', 'loss_cos'):
            assert rel(res[k], g[k]) < TOL_LOSS, (t, k, res[k], g[k])
        assert torch.equal(res['validity'], g['validity']) and torch.equal(res['sparse_depth'], g['sparse_depth'])
        assert nrel(res['output_depth'], g['output_depth']) < TOL_TENSOR, t
        assert nrel(res['emb'], g['emb'].float()) < 2e-3 and nrel(res['ref'], g['ref'].float()) < 2e-3      # stored as fp16
        if t == 0:
            gn = max(float(v.norm()) for v in fx['grads_step1'].values())
            for k in names:
                ref = fx['grads_step1'][k]
                # norm-wise, with an absolute floor for tensors whose gradient is rounding noise (conv bias before BN)
                assert float((res['grads'][k] - ref).norm()) < 2e-3 * float(ref.norm()) + 1e-5 * gn, k
    for k in names:
        # Adam's first steps move every element by ~lr*sign(g): elements whose gradient is rounding noise may differ by 2*lr*steps
        d = (sd[k] - fx['adapted'][k]).abs()
        assert float(d.max()) <= 2.001 * case['lr'] * case['steps'], k
    moved = [k for k in names if 'conv1_rgb_meta' in k or k.endswith('.1.weight') or k.endswith('bn1.weight')]
    for k in moved:
        assert nrel(sd[k], fx['adapted'][k]) < TOL_TENSOR, k
