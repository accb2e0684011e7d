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


def gather_shards(shard, n, rank, world, group=None):
    """COLLECTIVE: every rank's shard of a flat vector (ownership = shard_bounds) -> the full vector [n] on every rank."""
    begin, end = shard_bounds(n, rank, world)
    per = shard_bounds(n, 0, world)[1] if world > 1 else n
    mine = torch.zeros(max(per, 1), dtype=shard.dtype, device=shard.device)
    mine[:end - begin] = shard[:end - begin]
    if world == 1:
        return mine[:n]
    full = torch.empty(mine.numel() * world, dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


def adam_state_dict_from_flat(exp_avg, exp_avg_sq, step, shapes, param_groups):
    """Full flat Adam moments -> torch.optim.Adam.state_dict() layout over parameters of the given shapes (in flat order);
    `param_groups`: the optimiser's groups (their 'params' lists give the group sizes)."""
    state, off = {}, 0
    for i, shape in enumerate(shapes):
        n = 1
        for d in shape:
            n *= int(d)
        state[i] = {'step': torch.tensor(float(step)), 'exp_avg': exp_avg[off:off + n].reshape(shape).clone(),
                    'exp_avg_sq': exp_avg_sq[off:off + n].reshape(shape).clone()}
        off += n
    groups, idx = [], 0
    for g in param_groups:
        d = {k: v for k, v in g.items() if k != 'params'}
        d['params'] = list(range(idx, idx + len(g['params'])))
        idx += len(g['params'])
        groups.append(d)
    return {'state': state, 'param_groups': groups}


def flat_from_adam_state_dict(sd, n, device):
    """torch.optim.Adam.state_dict() -> (flat exp_avg [n], flat exp_avg_sq [n], step) or None for an empty state."""
    state = sd['state']
    if len(state) == 0:
        return None
    keys = sorted(state.keys())
    m = torch.cat([state[k]['exp_avg'].reshape(-1).to(device, torch.float32) for k in keys])
    v = torch.cat([state[k]['exp_avg_sq'].reshape(-1).to(device, torch.float32) for k in keys])
    if m.numel() != n:
        raise RuntimeError(f"optimiser state of a different model ({m.numel()} values, expected {n})")
    return m, v, int(float(state[keys[0]]['step']))


class PeerShardedAdam(torch.optim.Optimizer):
    """Gradient exchange + Adam for data-parallel training as ONE kernel over NVLink / NVSwitch peer memory
    (csrc/peer.cu, al_peer_adam_step) instead of `all_reduce(param.grad)` + the optimiser on every rank.

    All parameters and their gradients are moved into two flat buffers in symmetric memory
    (torch.distributed._symmetric_memory): `param.data` / `param.grad` become views, so the training kernels keep
    accumulating straight into the exchange buffer.  Per step: barrier -> every rank sums the gradients of ITS shard
    over all replicas (in-switch multimem.ld_reduce when the fabric has multicast, else peer loads), applies Adam
    with its shard of the moments, writes the new parameters into every replica -> barrier -> local memset of the
    gradient buffer.  Hyper-parameters as scripts/train.py:50-63 (`configure_optimizer`): weight decay on the MLP
    parameters only, mean over ranks folded into the kernel (grad_scale = 1 / world).

    A torch.optim.Optimizer (two parameter groups, 'encoding' and 'net', like the reference's Adam), so torch's
    lr schedulers attach to it; the kernel reads `param_groups[0]['lr']` every step.  `state_dict()` is a COLLECTIVE
    that returns the state in torch.optim.Adam's own format (moments all-gathered and cut per parameter), so a
    checkpoint written from N GPUs resumes on any other world size, on FusedAdam or on torch.optim.Adam;
    `load_state_dict()` takes that format and keeps this rank's shard.

    Install with `trainer.set_optimizer(PeerShardedAdam(model, ...))` (re-binds the lr scheduler)."""

    def __init__(self, model, lr=5e-3, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-6, group=None, use_multicast=None,
                 init_from=None):
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import call  # noqa: F401  (fails loudly without the library)
        if not dist.is_initialized():
            raise RuntimeError("PeerShardedAdam needs an initialised process group (parallel.init_distributed)")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        enc = [p for p in model.encoder.parameters() if p.requires_grad and p.numel() > 0]
        net = [p for p in model.network_parameters() if p.requires_grad and p.numel() > 0]
        groups = []
        if enc:
            groups.append({'name': 'encoding', 'params': enc, 'weight_decay': 0.0})
        groups.append({'name': 'net', 'params': net, 'weight_decay': weight_decay})
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0))
        self.params = enc + net                       # no-decay parameters first: one boundary, wd_begin
        self.sizes = [p.numel() for p in self.params]
        if any(n % 4 for n in self.sizes):
            raise RuntimeError("every parameter tensor must hold a multiple of 4 elements (16-byte aligned views)")
        self.n = sum(self.sizes)
        self.wd_begin = sum(p.numel() for p in enc)
        self.weight_decay = weight_decay
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
        for p, n in zip(self.params, self.sizes):
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
        self._arr = ctypes.c_void_p * self.world
        if init_from is not None:
            # continue from a replicated optimiser (FusedAdam / torch.optim.Adam over the same parameters): no collective
            self._load_full(init_from.state_dict())
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
             float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']), float(self.weight_decay),
             int(self.step_count), float(self.grad_scale), stream_ptr(dev))
        self.h_param.barrier(channel=1)               # every replica has its new parameters, every gradient has been read
        self.flat_grad.zero_()                        # local memset (HBM) instead of zeros over NVLink

    # ------------------------------------------------------------ checkpoints: torch.optim.Adam's format
    def state_dict(self):
        """COLLECTIVE (every rank must call it).  Same layout as torch.optim.Adam.state_dict() over
        [encoder params..., network params...]."""
        m = gather_shards(self.exp_avg, self.n, self.rank, self.world, self.group)
        v = gather_shards(self.exp_avg_sq, self.n, self.rank, self.world, self.group)
        return adam_state_dict_from_flat(m, v, self.step_count, [tuple(p.shape) for p in self.params], self.param_groups)

    def _load_full(self, sd):
        got = flat_from_adam_state_dict(sd, self.n, self.exp_avg.device)
        if got is None:
            return
        m, v, step = got
        self.exp_avg[:self.end - self.begin].copy_(m[self.begin:self.end])
        self.exp_avg_sq[:self.end - self.begin].copy_(v[self.begin:self.end])
        self.step_count = step
        for g, gs in zip(self.param_groups, sd.get('param_groups', [])):
            for k in ('lr', 'betas', 'eps', 'initial_lr'):
                if k in gs:
                    g[k] = gs[k]

    def load_state_dict(self, sd):
        """Takes torch.optim.Adam's format (what state_dict() returns, whatever the world size it was saved from)."""
        self._load_full(sd)


def broadcast_parameters(model, src=0, group=None):
    """Make replicas start identical (parameters and occupancy buffers)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_frames(n_frames, rank, world):
    """Frame i -> rank i mod world (export / render: no communication)."""
    return list(range(rank, n_frames, world))
