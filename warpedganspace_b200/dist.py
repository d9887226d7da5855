"""One-process-per-GPU plumbing (torch.distributed): the paired step shards over latents, so the only
collective is the sum of the flat S / R gradient buffers.  Device-agnostic (NCCL on GPUs, gloo in the
CPU tests); replaces the reference's single-process nn.DataParallel (lib/trainer.py:16-21,162-166)."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment. Returns (world, rank, local_rank)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return world, rank, local


def shard_range(n, rank, world, require_equal=False):
    """Contiguous [lo, hi) slice of n latents owned by `rank` (sizes differ by at most one).  Training passes
    require_equal=True: each rank's loss is the mean over its own shard and the gradients are summed and scaled by
    1/world, which equals the reference's global batch mean (lib/trainer.py:245-249) only for equal, non-empty shards."""
    if require_equal and (n % world != 0 or n < world):
        raise ValueError('batch size %d does not split into %d equal non-empty shards: the mean of per-rank losses would '
                         'not be the batch mean (pick a batch size that is a multiple of the number of ranks)' % (n, world))
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(tensors, group=None):
    """In-place sum over ranks of each tensor (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def is_parallel(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def broadcast_(tensors, src=0, group=None):
    """In-place broadcast of each tensor from rank `src` (no-op without a process group)."""
    if is_parallel(group):
        for t in tensors:
            dist.broadcast(t, src=src, group=group)
    return tensors


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
