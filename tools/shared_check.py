#!/usr/bin/env python
"""Shared-model mode check, launched as `python -m torch.distributed.run --nproc-per-node W tools/shared_check.py` (W >= 2 GPUs of one node).

Every rank adapts the SAME model on its own batch-1 frame per step through the peer-memory path (SyncBatchNorm sums + gradient all-reduce
fused with Adam, csrc/peer_comm.cuh).  Checked: (1) the adapted tensors, BatchNorm buffers and Adam moments are BIT-IDENTICAL on all
ranks after every step; (2) they equal a single-GPU step on the batch of all W frames (what DDP + SyncBatchNorm is defined to reproduce)
up to fp32 summation order; (3) the CUDA-graph replay of the step gives the same result as eager launches.  Prints one JSON line.
`--stage init | head` runs the same three checks on the source-domain preparation steps (the reference's src/init_main.py and
src/head_main.py are DDP + SyncBatchNorm trainers: one model, every rank its own batch, gradients averaged)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from tta_depth_completion_b200 import ExternalModel_Adapt, sharding, synthetic

MODE, CAP, LR = 'meta_selfsup_seq_2layers_ema', 80.0, 1e-4
H, W_, STEPS = 64, 128, 3


STAGE = 'tta'
for _i, _a in enumerate(sys.argv):
    if _a == '--stage':
        STAGE = sys.argv[_i + 1]


def run_step(model, image, sparse, dense, graph):
    if STAGE == 'tta':
        sharding.shared_model_step(model, image, sparse, LR, 1.0, 1.0, 0.1, graph=graph)
    elif STAGE == 'init':
        model.init_step(image, sparse, dense, 1e-3, graph=graph)
    else:
        model.head_step(image, sparse, 1e-3, graph=graph)


def make_model(dev, sd):
    m = ExternalModel_Adapt('msg_chn', 0.0, 100.0, max_input_depth=CAP, device=dev)
    # the kernel dispatch depends on the pixel count of a map (N x H x W): force the tcgen05 kernels everywhere, so that the per-rank
    # batch-1 engines and the single-GPU batch-W engine run the SAME kernels and differ by summation order only
    m.model.engine_options = {'tc_min_pixels': 0, 'tc_s2_min_pixels': 0, 'tc_t2_min_pixels': 0, 'tc_head_min_pixels': 0, 'tc_stem_min_pixels': 0}
    m._prepare_head(MODE)
    m.load_state_dict(sd)
    if STAGE == 'head':
        torch.manual_seed(11)                                   # the same fresh heads on every rank (src/head_main.py:268)
        m.prepare_parameters('head_selfsup_ema')
    m.set_image_normalization((1 / 255.0,) * 3, (0.0,) * 3)
    m.train()
    return m


def nrel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    sd = synthetic.get_checkpoint('kitti_2layers_a', MODE) if synthetic.fitted_checkpoint_available('kitti_2layers_a') else synthetic.make_synthetic_checkpoint(0, MODE)
    out = {'world': world}
    torch.cuda.set_stream(torch.cuda.Stream(dev))          # a step can only be captured on a non-default stream
    for use_graph in (False, True):
        model = make_model(dev, sd)
        comm = sharding.enable_shared_model(model)
        losses = []
        for t in range(STEPS):
            image, sparse, dense = synthetic.synthetic_frame(40, t * world + rank, 1, H, W_, 'kitti')
            run_step(model, image.to(dev), sparse.to(dev), dense.to(dev), use_graph)
            losses.append(model.last_losses()['loss'])
            # (1) replicas identical, bit for bit
            flat = torch.cat([model.model._flat['param'], model.model._flat['m'], model.model._flat['v'], model.model._flat['grad']])
            gathered = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            for r in range(world):
                assert torch.equal(gathered[r], gathered[0]), 'replicas differ at step %d (rank %d vs 0, graph=%s)' % (t, r, use_graph)
            bufs = torch.cat([v.float().flatten() for k, v in model.state_dict().items() if 'running' in k])
            gb = [torch.empty_like(bufs) for _ in range(world)]
            dist.all_gather(gb, bufs)
            for r in range(world):
                assert torch.equal(gb[r], gb[0]), 'BatchNorm buffers differ at step %d' % t
        assert comm.error() == 0, 'peer exchange %d timed out' % (comm.error() - 1)
        lt = torch.tensor(losses, device=dev, dtype=torch.float64)
        dist.all_reduce(lt)
        mean_losses = (lt / world).tolist()
        key = 'graph' if use_graph else 'eager'
        out[key] = {'mean_loss': mean_losses}
        # (2) single-GPU step on the whole batch (rank 0 only)
        if rank == 0:
            big = make_model(dev, sd)
            big_losses = []
            for t in range(STEPS):
                frames = [synthetic.synthetic_frame(40, t * world + r, 1, H, W_, 'kitti') for r in range(world)]
                image = torch.cat([f[0] for f in frames]).to(dev)
                sparse = torch.cat([f[1] for f in frames]).to(dev)
                dense = torch.cat([f[2] for f in frames]).to(dev)
                if STAGE == 'tta':
                    big.tta_step(image, sparse, LR, 1.0, 1.0, 0.1)
                elif STAGE == 'init':
                    big.init_step(image, sparse, dense, 1e-3)
                else:
                    big.head_step(image, sparse, 1e-3)
                big_losses.append(big.last_losses()['loss'])
            sa, sb = model.state_dict(), big.state_dict()
            errs = {}
            for k in model.model._adapt_names:
                if k in ('conv1_rgb_meta.conv1_meta.1.bias', 'pred.0.bias'):     # bias in front of a train-mode BatchNorm: analytically zero gradient, pure rounding noise
                    continue
                upd = nrel(sd[k].to(dev), sb[k])
                errs[k] = (nrel(sa[k], sb[k]), upd)
            worst = max(e / max(u, 1e-30) for e, u in errs.values())
            buf_err = max(nrel(sa[k].float(), sb[k].float()) for k in sa if 'running' in k)
            loss_err = max(abs(a - b) / abs(b) for a, b in zip(mean_losses, big_losses))
            out[key].update(worst_error_over_update=worst, bn_buffer_error=buf_err, loss_error=loss_err, big_batch_loss=big_losses)
            assert loss_err < 2e-3, (mean_losses, big_losses)
            assert buf_err < 1e-4, buf_err
            assert worst < 0.05, errs           # fp32 summation order only (per-rank wgrad partial sums vs one sum) through Adam's 1/sqrt(v)
        dist.barrier()
    out['stage'] = STAGE
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
