"""Shared-model mode (BASELINE.json configs[4]) on real GPUs: SyncBatchNorm sums and the gradient all-reduce fused with Adam through NVLink
peer memory (csrc/peer_comm.cuh).  The world-2 check needs two GPUs (tools/shared_check.py under torch.distributed.run); with one GPU only
the world-1 plumbing runs (communicator block, engine binding, same result as the plain step)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world1_communicator_is_a_no_op():
    from oracle import msgchn_oracle as O
    from test_msgchn_step_gpu import make_model
    from golden_util import W_SD, W_SM, W_COS
    from tta_depth_completion_b200 import sharding
    mode, cap = 'meta_selfsup_seq_2layers_ema', 80.0
    sd = O.make_synthetic_checkpoint(0, mode)
    a, b = make_model(mode, sd, cap), make_model(mode, sd, cap)
    comm = sharding.enable_shared_model(b)
    assert comm.world == 1
    for t in range(2):
        image, sparse, _ = O.synthetic_frame(3, t, 1, 64, 128, 'kitti')
        a.tta_step(image.cuda(), sparse.cuda(), 1e-4, W_SD, W_SM, W_COS)
        sharding.shared_model_step(b, image.cuda(), sparse.cuda(), 1e-4, W_SD, W_SM, W_COS)
    assert comm.error() == 0
    # same kernels, same inputs -> same weights
    for k in a.model._adapt_names:
        assert torch.equal(a.state_dict()[k], b.state_dict()[k]), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs of one node')
def test_world2_replicas_identical_and_equal_to_the_big_batch():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', '29731',
           os.path.join(ROOT, 'tools', 'shared_check.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('{')][-1]
    out = json.loads(line)
    assert out['world'] == 2 and 'worst_error_over_update' in out['eager'] and 'worst_error_over_update' in out['graph']
