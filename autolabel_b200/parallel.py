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
import ctypes
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


def shard_bounds(n, rank, world, align=4):
    """Element range [begin, end) of a flat vector of n elements owned by `rank`: equal shares rounded up to `align`
    elements, the last ranks may own less (or nothing)."""
    per = -(-n // world)
    per = -(-per // align) * align
    begin = min(rank * per, n)
    return begin, min(begin + per, n)


class PeerShardedAdam:
    """Gradient exchange + Adam for data-parallel training as ONE kernel over NVLink / NVSwitch peer memory
    (csrc/peer.cu, al_peer_adam_step) instead of `all_reduce(param.grad)` + the optimiser on every rank.

    All parameters and their gradients are moved into two flat buffers in symmetric memory
    (torch.distributed._symmetric_memory): `param.data` / `param.grad` become views, so the training kernels keep
    accumulating straight into the exchange buffer.  Per step: barrier -> every rank sums the gradients of ITS shard
    over all replicas (in-switch multimem.ld_reduce when the fabric has multicast, else peer loads), applies Adam
    with its shard of the moments, writes the new parameters into every replica -> barrier -> local memset of the
    gradient buffer.  Hyper-parameters as scripts/train.py:50-63 (`configure_optimizer`): weight decay on the MLP
    parameters only, mean over ranks folded into the kernel (grad_scale = 1 / world).

    Drop-in for the trainer: `trainer.optimizer = trainer.optimizers[0] = PeerShardedAdam(model, ...)`,
    `trainer.grad_sync = None`."""

    def __init__(self, model, lr=5e-3, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-6, group=None, use_multicast=None):
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import call  # noqa: F401  (fails loudly without the library)
        if not dist.is_initialized():
            raise RuntimeError("PeerShardedAdam needs an initialised process group (parallel.init_distributed)")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        enc = [p for p in model.encoder.parameters() if p.requires_grad and p.numel() > 0]
        net = [p for p in model.network_parameters() if p.requires_grad and p.numel() > 0]
        self.params = enc + net                       # no-decay parameters first: one boundary, wd_begin
        sizes = [p.numel() for p in self.params]
        if any(n % 4 for n in sizes):
            raise RuntimeError("every parameter tensor must hold a multiple of 4 elements (16-byte aligned views)")
        self.n = sum(sizes)
        self.wd_begin = sum(p.numel() for p in enc)
        dev = self.params[0].device
        if hasattr(symm_mem, "enable_symm_mem_for_group"):
            try:
                symm_mem.enable_symm_mem_for_group(self.group.group_name)
            except Exception:
                pass
        self.flat_param = symm_mem.empty(self.n, dtype=torch.float32, device=dev)
        self.flat_grad = symm_mem.empty(self.n, dtype=torch.float32, device=dev)
        self.flat_grad.zero_()
        off = 0
        for p, n in zip(self.params, sizes):
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            g = self.flat_grad[off:off + n].view_as(p)
            if p.grad is not None:
                g.copy_(p.grad)
            p.grad = g
            off += n
        self.h_param = symm_mem.rendezvous(self.flat_param, self.group)
        self.h_grad = symm_mem.rendezvous(self.flat_grad, self.group)
        self.param_ptrs = [int(x) for x in self.h_param.buffer_ptrs]
        self.grad_ptrs = [int(x) for x in self.h_grad.buffer_ptrs]
        mc_p = int(getattr(self.h_param, "multicast_ptr", 0) or 0)
        mc_g = int(getattr(self.h_grad, "multicast_ptr", 0) or 0)
        if use_multicast is None:
            # measured (profiles/r1h_exchange_*gpu.json): plain peer loads win at 2 GPUs (0.12 vs 0.20 ms), the two forms
            # tie at 4 (0.18 ms); the multicast form moves 1/W of the bytes per rank, so it is the default from 4 GPUs on
            use_multicast = self.world >= 4
        self.multicast = bool(use_multicast and mc_p and mc_g)
        self.mc_param, self.mc_grad = (mc_p, mc_g) if self.multicast else (None, None)
        self.begin, self.end = shard_bounds(self.n, self.rank, self.world)
        self.exp_avg = torch.zeros(max(self.end - self.begin, 4), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.step_count = 0
        self.grad_scale = 1.0 / self.world
        self.param_groups = [{'lr': lr, 'betas': betas, 'eps': eps, 'weight_decay': weight_decay, 'params': self.params}]
        self._arr = ctypes.c_void_p * self.world
        torch.cuda.synchronize(dev)
        self.h_param.barrier(channel=0)               # every replica has been filled before anyone reads or writes it
        torch.cuda.synchronize(dev)

    def zero_grad(self, set_to_none=False):
        """Gradients are cleared inside step(), after the closing barrier."""
        return

    @torch.no_grad()
    def step(self, closure=None):
        from ._lib import call, ptr, stream_ptr
        g = self.param_groups[0]
        dev = self.flat_param.device
        self.step_count += 1
        self.h_grad.barrier(channel=0)                # all ranks have finished their backward pass
        call("al_peer_adam_step", self._arr(*self.grad_ptrs), self._arr(*self.param_ptrs), self.mc_grad, self.mc_param,
             ptr(self.exp_avg), ptr(self.exp_avg_sq), self.begin, self.end, self.wd_begin, self.world, self.rank,
             float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']), float(g['weight_decay']),
             int(self.step_count), float(self.grad_scale), stream_ptr(dev))
        self.h_param.barrier(channel=1)               # every replica has its new parameters, every gradient has been read
        self.flat_grad.zero_()                        # local memset (HBM) instead of zeros over NVLink

    def state_dict(self):
        return {'step': self.step_count, 'begin': self.begin, 'end': self.end, 'exp_avg': self.exp_avg,
                'exp_avg_sq': self.exp_avg_sq, 'lr': self.param_groups[0]['lr']}

    def load_state_dict(self, sd):
        if (sd['begin'], sd['end']) != (self.begin, self.end):
            raise RuntimeError("optimiser shard of a different world size / rank")
        self.step_count = int(sd['step'])
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        self.param_groups[0]['lr'] = sd.get('lr', self.param_groups[0]['lr'])


def broadcast_parameters(model, src=0, group=None):
    """Make replicas start identical (parameters and occupancy buffers)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_frames(n_frames, rank, world):
    """Frame i -> rank i mod world (export / render: no communication)."""
    return list(range(rank, n_frames, world))
