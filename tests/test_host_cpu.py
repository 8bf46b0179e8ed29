"""CPU-side checks: the C-ABI library loads and exports every symbol include/ptta_b200.h declares, the engine's host logic
(no kernels) behaves, the facade builds the reference's state dict, and the sharding helpers work under a world-size-2
gloo group.  No compute call is made here (no GPU in this tier)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import msgchn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from tta_depth_completion_b200.build import build_library
    build_library()
    from tta_depth_completion_b200 import _lib
    return _lib


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, 'include', 'ptta_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = sorted(set(re.findall(r'\b(ptta_\w+)\s*\(', header)))
    assert len(declared) >= 35
    assert sorted(lib.PROTOTYPES) == declared, 'prototype parser and header disagree'
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.ptta_version() >= 100
    assert lib.last_error() == ''


@pytest.mark.parametrize('mode,n_adapt', [('meta_selfsup_seq_2layers_ema', 7), ('meta_selfsup_seq_1layer_ema', 2)])
def test_engine_host_logic_without_gpu(lib, mode, n_adapt):
    L = lib.lib()
    h = ctypes.c_void_p()
    assert L.ptta_msgchn_create(ctypes.byref(h), 1, 352, 1216, mode.encode()) == 0, lib.last_error()
    try:
        nbytes = L.ptta_msgchn_workspace_bytes(h)
        assert 1 << 29 < nbytes < 1 << 32            # ~1-2 GB of activations at KITTI size, N=1
        keys = [L.ptta_msgchn_key(h, i).decode() for i in range(L.ptta_msgchn_num_keys(h))]
        sd = O.make_synthetic_checkpoint(0, mode)
        assert set(keys) <= set(sd), sorted(set(keys) - set(sd))[:5]
        # every parameter and BatchNorm buffer of the checkpoint, except the EMA copy's buffers (stage 2 updates proj_t's PARAMETERS only)
        assert {k for k in sd if not (k.startswith('proj_t.') and not k.endswith(('.weight', '.bias')))} == set(keys)
        # unbound engine refuses to run, loudly
        assert L.ptta_msgchn_pack_weights(h, None) != 0
        assert 'workspace' in lib.last_error()
    finally:
        L.ptta_msgchn_destroy(h)


def test_engine_rejects_bad_arguments(lib):
    L = lib.lib()
    h = ctypes.c_void_p()
    # any H, W is accepted since the engine pads to /16 itself (a4); degenerate shapes are not
    assert L.ptta_msgchn_create(ctypes.byref(h), 1, 0, 1216, b'meta_selfsup_seq_2layers_ema') != 0
    assert 'bad shape' in lib.last_error()
    # general-channel conv family: channel counts must be multiples of 64, the 1x1/s2 data gradient only exists folded
    assert L.ptta_convg_packed_elems(0, 0, 48, 0, 64, 0) < 0
    assert 'multiples of 64' in lib.last_error()
    assert L.ptta_convg_packed_elems(3, 1, 64, 0, 128, 0) < 0
    assert L.ptta_convg_packed_elems(0, 0, 64, 64, 64, 0) == 18 * 64 * 64          # 9 taps x 2 sources x [64][64]
    assert L.ptta_convg_packed_elems(1, 1, 64, 0, 128, 1) == (9 * 2 + 2) * 64 * 64   # 3x3/s2 dgrad (K = 128) + folded 1x1 shortcut
    assert L.ptta_msgchn_create(ctypes.byref(h), 1, 352, 1216, b'selfsup_only') != 0
    assert 'meta' in lib.last_error()
    assert L.ptta_msgchn_create(ctypes.byref(h), 0, 352, 1216, b'meta_selfsup_seq_2layers_ema') != 0
    assert L.ptta_gemm_bf16(None, None, None, None, 16, 48, 40, None) != 0          # K not a multiple of 32
    assert 'multiple of 32' in lib.last_error()


@pytest.mark.parametrize('mode', ['meta_selfsup_seq_2layers_ema', 'meta_selfsup_seq_1layer_ema'])
def test_facade_state_matches_reference_keys(mode):
    from tta_depth_completion_b200.external_model_adapt import build_base_state, add_head_state
    sd = add_head_state(build_base_state(), mode)
    ref = O.make_synthetic_checkpoint(0, mode)
    assert list(sd.keys()) == list(ref.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape) and sd[k].dtype == ref[k].dtype, k
    if '2layers' in mode:
        assert len(sd) == 132


def test_facade_refuses_cpu_and_unknown_models():
    from tta_depth_completion_b200 import ExternalModel_Adapt
    with pytest.raises(ValueError):
        ExternalModel_Adapt('resnet', 0.0, 100.0, device=torch.device('cpu'))
    with pytest.raises(NotImplementedError):
        ExternalModel_Adapt('costdcnet', 0.0, 100.0, device=torch.device('cpu'))
    nl = ExternalModel_Adapt('nlspn', 0.0, 100.0, device=torch.device('cpu'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        nl._prepare_head('meta_selfsup_seq_1layer_ema')
    m = ExternalModel_Adapt('msg_chn', 0.0, 100.0, device=torch.device('cpu'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m._prepare_head('meta_selfsup_seq_2layers_ema')


def test_shard_assignment():
    from tta_depth_completion_b200.sharding import shard_sequences, shard_sizes
    for n, w in [(8, 8), (10, 4), (3, 8), (0, 2)]:
        shards = [shard_sequences(n, w, r) for r in range(w)]
        assert sorted(sum(shards, [])) == list(range(n))
        assert [len(s) for s in shards] == shard_sizes(n, w)
        assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    with pytest.raises(ValueError):
        shard_sequences(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from tta_depth_completion_b200.sharding import allreduce_mean_, shard_sequences
rank, world = int(sys.argv[2]), 2
os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group('gloo', rank=rank, world_size=world)
g = torch.Generator().manual_seed(100 + rank)
grad = torch.randn(74080, generator=g)                       # flat adapted-gradient buffer of MSG-CHN `2layers`
want = (torch.randn(74080, generator=torch.Generator().manual_seed(100)) + torch.randn(74080, generator=torch.Generator().manual_seed(101))) / 2
allreduce_mean_(grad)
assert torch.allclose(grad, want, atol=1e-7), float((grad - want).abs().max())
# identical update on both ranks -> identical parameters
p = torch.ones(74080); p -= 1e-4 * grad.sign()
gathered = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(gathered, p)
assert torch.equal(gathered[0], gathered[1])
seqs = shard_sequences(5, world, rank)
out = [None, None]
dist.all_gather_object(out, seqs)
assert sorted(out[0] + out[1]) == [0, 1, 2, 3, 4]
dist.destroy_process_group()
print('ok', rank)
'''


def test_shared_model_gradient_allreduce_gloo_world2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and 'ok %d' % r in o, o


def test_nlspn_facade_state_matches_reference_keys():
    """key set, order and shapes of the NLSPN facade's fresh state == the reference's NLSPNModel_Adapt.state_dict() (manifest written
    from the live reference by oracle/gen_golden_nlspn_net.py)"""
    import json
    import os
    from tta_depth_completion_b200.nlspn_model_adapt import build_nlspn_state
    from tta_depth_completion_b200.nlspn_engine import adapt_parameter_names
    from oracle import nlspn_oracle as NO
    sd = build_nlspn_state('meta_selfsup_seq_1layer_ema')
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, '..', 'oracle', 'nlspn_state_manifest.json')) as f:
        manifest = json.load(f)
    assert set(sd) == set(manifest)
    for k, shape in manifest.items():
        assert list(sd[k].shape) == shape, k
    ref = NO.make_synthetic_checkpoint(0)
    assert list(sd.keys()) == list(ref.keys())
    assert adapt_parameter_names(sd) == NO.adapt_parameter_names(ref, 'meta_bn')
    with pytest.raises(NotImplementedError):
        build_nlspn_state('meta_selfsup_seq_2layers_ema')
