"""Multi-GPU host logic for the TTA path (one process per GPU).

Default mode ("shards"): continual adaptation is a serial recurrence inside a sequence but sequences are independent,
so rank r adapts its own model copy on sequences r, r+W, r+2W, ... with NO data-path collective (the reference instead
runs shared-model DDP + SyncBN + one barrier per step: src/tta_main.py:101-111,354,804).

Optional mode ("shared"): every rank holds the same model, processes its own batch, and the adapted-parameter gradients
(one flat fp32 buffer, 74 080 floats = 296 KB for MSG-CHN `2layers`) are mean-all-reduced before the fused Adam step, so
all replicas apply the identical update (the reference's DDP all-reduces all 1.53 M gradients)."""
import torch
import torch.distributed as dist


def shard_sequences(n_sequences, world_size, rank):
    """Sequence ids adapted by `rank` (strided like DistributedSampler, src/tta_main.py:17-21, without padding: a rank
    may get one sequence fewer)."""
    if not (0 <= rank < world_size):
        raise ValueError('rank %d outside world of %d' % (rank, world_size))
    return list(range(rank, n_sequences, world_size))


def shard_sizes(n_sequences, world_size):
    return [len(range(r, n_sequences, world_size)) for r in range(world_size)]


def allreduce_mean_(flat, group=None):
    """In-place mean all-reduce of one flat buffer (NCCL on CUDA tensors, gloo on CPU tensors)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    return flat


def shared_model_step(model, image_raw, sparse_depth, learning_rate, w_sd=1.0, w_sm=1.0, w_cos=0.1, group=None):
    """One shared-model adaptation step: local forward/loss/backward in the engine, mean all-reduce of the flat
    adapted-gradient buffer, fused Adam (identical on every rank)."""
    wrapper = model.model
    eng = wrapper._engine_for(image_raw)
    hyper = (learning_rate, (0.9, 0.999), 1e-8, 0.0)
    if getattr(eng, '_hyper', None) != hyper:
        eng.set_adam(learning_rate, step_count=-1)
        eng._hyper = hyper
    from . import ops
    d_f, v_f = ops.outlier_removal(sparse_depth.contiguous())
    eng.forward(image_raw, d_f, model.max_input_depth, True, wrapper.img_scale, wrapper.img_shift)
    eng.loss(image_raw, d_f, v_f, model.max_input_depth, w_sd, w_sm, w_cos)
    eng.backward(1.0)
    allreduce_mean_(wrapper._flat['grad'], group)
    eng.adam_step()
    model._last_engine = eng
