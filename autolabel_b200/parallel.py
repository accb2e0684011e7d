"""Ray-sharded data-parallel training: one process per GPU, every rank marches its own
ray batch; the only exchange per step is ONE sum all-reduce of the parameter gradients (hash table
+ four MLPs) over NCCL / NVLink, issued on the gradient buffers the kernels accumulated into (no
staging copy).  The mean over ranks is folded into the fused Adam step (grad_scale = 1 / world).
The occupancy grid stays replica-identical because its refresh draws from the (identically seeded)
global torch RNG and queries identical parameters; each rank's ray sampler uses its own generator.
Full-frame rendering shards by frame with no communication (`shard_frames`).

The reference has no equivalent: it only wraps the model in DistributedDataParallel when
world_size > 1 and never launches more than one process (torch_ngp/nerf/utils.py:301-304,
scripts/train.py:80-92).
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


class GradientAllReduce:
    """Callable plugged into SimpleTrainer.grad_sync: sum-all-reduce every parameter gradient."""

    def __init__(self, params, optimizer=None, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if optimizer is not None and hasattr(optimizer, "grad_scale"):
            optimizer.grad_scale = 1.0 / self.world      # mean over ranks, applied inside the Adam kernel
            self.average_here = False
        else:
            self.average_here = True

    def __call__(self):
        if self.world == 1:
            return
        handles = []
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)      # every rank must join the collective
            handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for h in handles:
            h.wait()
        if self.average_here:
            for p in self.params:
                p.grad.div_(self.world)


def broadcast_parameters(model, src=0, group=None):
    """Make replicas start identical (parameters and occupancy buffers)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_frames(n_frames, rank, world):
    """Frame i -> rank i mod world (export / render: no communication)."""
    return list(range(rank, n_frames, world))
