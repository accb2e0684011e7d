"""Training loop — the interface of the reference's ``autolabel/trainer.py`` (``SimpleTrainer``
:14-147) for the B200 path.

Same losses as ``train_step`` (:54-94): ``rgb_w * MSE + depth_w * mean|d - d_gt| over d_gt > 0.01 +
feature_w * L1(features[:, :F_gt]) + sem_w * CE(logits[sem >= 0])``, same Adam configuration as
``scripts/train.py:50-63`` (``configure_optimizer``).  Differences:
  * the occupancy grid is refreshed every 16 steps (``update_extra_state``; upstream does it in
    ``Trainer.train_one_epoch``, torch_ngp/nerf/utils.py:906-910 — autolabel's loop never does,
    which would leave the bitfield empty with cuda_ray=True);
  * the step is free of host synchronisation: masked means instead of boolean-mask indexing, no
    ``.item()`` on the loss (the progress string is refreshed every ``log_interval`` steps);
  * no GradScaler is needed: the kernels scale fp16 gradients internally (csrc/mlp.cu) and deliver
    fp32 gradients; ``fp16=True`` is accepted for interface compatibility.
  * ``opt.train_t_thresh`` (optional, default 0 = the reference's semantics: every marched sample is composited)
    opts into training-time early termination (renderer.NeRFRenderer.train_t_thresh).
Checkpoints carry the reference Trainer's key set (torch_ngp/nerf/utils.py:1124-1260: epoch, global_step, stats,
precision, mean_count, mean_density, opt0.., lr_sched0.., scaler, ema, model) in torch.optim.Adam's own optimiser-state
format, so they move between this trainer, the reference trainer and any world size.
"""
import glob
import os

import torch
import torch.nn.functional as F

from .optim import FusedAdam

DEPTH_EPSILON = 0.01


def configure_optimizer(model, lr=5e-3, weight_decay=1e-6):
    """scripts/train.py:50-63: encoder parameters without weight decay, MLPs with 1e-6."""
    groups = []
    enc = list(model.encoder.parameters())
    if enc:
        groups.append({'name': 'encoding', 'params': enc})
    groups.append({'name': 'net', 'params': model.network_parameters(), 'weight_decay': weight_decay})
    return FusedAdam(groups, lr=lr, betas=(0.9, 0.99), eps=1e-15)


BATCH_KEYS = ('rays_o', 'rays_d', 'direction_norms', 'pixels', 'depth', 'features', 'semantic')


def batch_layout(n_rays, feature_dim):
    """Byte layout of a training batch (the dict of autolabel/dataset.py:232-241) packed into ONE flat buffer, so a
    step's inputs move with a single copy: {key: (byte offset, shape, dtype)}; int64 labels last (8-byte aligned)."""
    lay, off = {}, 0
    for k, shape in (('rays_o', (n_rays, 3)), ('rays_d', (n_rays, 3)), ('direction_norms', (n_rays, 1)),
                     ('pixels', (n_rays, 3)), ('depth', (n_rays,)), ('features', (n_rays, feature_dim))):
        if k == 'features' and feature_dim == 0:
            continue
        numel = 1
        for d in shape:
            numel *= d
        lay[k] = (off, shape, torch.float32)
        off += 4 * numel
    off = (off + 7) // 8 * 8
    lay['semantic'] = (off, (n_rays,), torch.int64)
    return lay, off + 8 * n_rays


class PackedBatch(dict):
    """A training batch whose tensors are views of one flat uint8 buffer (`.flat`): `SimpleTrainer` moves it with a
    single copy.  Behaves as the plain dict everywhere else."""

    def __init__(self, n_rays, feature_dim, device='cpu', pin=False):
        super().__init__()
        self.layout, nbytes = batch_layout(n_rays, feature_dim)
        self.shape_key = (n_rays, feature_dim)
        self.flat = torch.empty(nbytes, dtype=torch.uint8, device=device, pin_memory=bool(pin and str(device) == 'cpu'))
        for k, (off, shape, dtype) in self.layout.items():
            numel = 1
            for d in shape:
                numel *= d
            self[k] = self.flat[off:off + numel * dtype.itemsize].view(dtype).view(*shape)

    @classmethod
    def pack(cls, data, device=None, pin=False):
        """Copy a batch dict into a packed one (on `device`, default: where the rays live)."""
        n = data['rays_o'].reshape(-1, 3).shape[0]
        fd = data['features'].shape[-1] if 'features' in data else 0
        out = cls(n, fd, device if device is not None else data['rays_o'].device, pin)
        for k in out.layout:
            out[k].copy_(data[k].reshape(out[k].shape))
        return out


class _EMA:
    """Minimal torch_ema.ExponentialMovingAverage (update / copy_to / store / restore / state_dict)."""

    def __init__(self, params, decay):
        self.params = [p for p in params if p.requires_grad]
        self.decay = decay
        self.shadow = [p.detach().clone() for p in self.params]
        self.backup = None
        self.num_updates = 0

    @torch.no_grad()
    def update(self):
        self.num_updates += 1
        d = min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))
        torch._foreach_mul_(self.shadow, d)
        torch._foreach_add_(self.shadow, [p.detach() for p in self.params], alpha=1 - d)

    @torch.no_grad()
    def store(self):
        self.backup = [p.detach().clone() for p in self.params]

    @torch.no_grad()
    def copy_to(self):
        for p, s in zip(self.params, self.shadow):
            p.copy_(s)

    @torch.no_grad()
    def restore(self):
        for p, b in zip(self.params, self.backup):
            p.copy_(b)
        self.backup = None

    def state_dict(self):
        """torch_ema's key set (decay, num_updates, shadow_params, collected_params), so a reference checkpoint's
        'ema' entry (torch_ngp/nerf/utils.py:1150-1151) loads here and vice versa."""
        return {'decay': self.decay, 'num_updates': self.num_updates, 'shadow_params': self.shadow,
                'collected_params': self.backup}

    def load_state_dict(self, sd):
        self.decay, self.num_updates = sd['decay'], sd['num_updates']
        for s, t in zip(self.shadow, sd.get('shadow_params', sd.get('shadow'))):
            s.copy_(t)


class SimpleTrainer:

    def __init__(self, name, opt, model, optimizer=None, lr_scheduler=None, criterion=None, device='cuda:0',
                 fp16=True, ema_decay=None, workspace=None, use_checkpoint='latest', update_interval=16,
                 log_interval=100, max_keep_ckpt=2, world_size=1, local_rank=0, **kwargs):
        self.name, self.opt, self.device = name, opt, torch.device(device)
        self.model = model.to(self.device)
        self.fp16 = fp16
        self.world_size, self.local_rank = world_size, local_rank
        if getattr(opt, 'train_t_thresh', None) is not None:
            self.model.train_t_thresh = float(opt.train_t_thresh)      # opt-in early termination (default 0: exact)
        self.optimizer = optimizer(self.model) if callable(optimizer) else (optimizer or configure_optimizer(self.model, getattr(opt, 'lr', 5e-3)))
        self.optimizers = [self.optimizer]
        self._lr_scheduler_factory = lr_scheduler if callable(lr_scheduler) else None
        self.lr_scheduler = lr_scheduler(self.optimizer) if callable(lr_scheduler) else lr_scheduler
        self.stats = {"loss": [], "valid_loss": [], "results": [], "checkpoints": [], "best_result": None}
        self.criterion = criterion or torch.nn.MSELoss(reduction='none')
        self.ema = _EMA(self.model.parameters(), ema_decay) if ema_decay is not None else None
        self.workspace = workspace
        self.max_keep_ckpt = max_keep_ckpt
        self.update_interval = update_interval
        self.log_interval = log_interval
        self.epoch = 0
        self.global_step = 0
        self.last_loss = None
        self.grad_sync = None      # set by parallel.DataParallel: called between backward and step
        self.fused_step = kwargs.get('fused_step', True)
        # replay the fused step as a CUDA graph (AL_NO_GRAPH=1 or use_graph=False: launch kernel by kernel)
        self.use_graph = kwargs.get('use_graph', os.environ.get('AL_NO_GRAPH', '0') != '1')
        self._graph_state = None
        self.graph_kernel_launches = 0
        self.last_loss_parts = None
        if workspace is not None:
            os.makedirs(os.path.join(workspace, 'checkpoints'), exist_ok=True)
            if use_checkpoint == 'latest':
                self.load_checkpoint()

    @property
    def lr_schedulers(self):
        return [] if self.lr_scheduler is None else [self.lr_scheduler]

    def set_optimizer(self, optimizer):
        """Swap the optimiser after construction (e.g. parallel.PeerShardedAdam, which needs the process group and the
        parameters first).  The lr scheduler is rebuilt on the new optimiser from the factory given to the constructor
        and continues from the old one's state, so the schedule keeps reaching the kernel."""
        old = self.lr_scheduler
        self.optimizer = optimizer
        self.optimizers = [optimizer]
        self._graph_state = None
        if old is not None:
            if self._lr_scheduler_factory is None:
                raise RuntimeError("set_optimizer: pass lr_scheduler as a factory (callable) so it can be rebuilt on the "
                                   "new optimiser")
            self.lr_scheduler = self._lr_scheduler_factory(optimizer)
            self.lr_scheduler.load_state_dict({k: v for k, v in old.state_dict().items() if k != '_last_lr'})
            for g, lr in zip(optimizer.param_groups, old.get_last_lr()):
                g['lr'] = lr
        return optimizer

    # ------------------------------------------------------------ steps
    def train_step(self, data):
        dev = self.device
        rays_o = data['rays_o'].to(dev, non_blocking=True)
        rays_d = data['rays_d'].to(dev, non_blocking=True)
        direction_norms = data['direction_norms'].to(dev, non_blocking=True)
        gt_rgb = data['pixels'].to(dev, non_blocking=True)
        gt_depth = data['depth'].to(dev, non_blocking=True)
        gt_semantic = data['semantic'].to(dev, non_blocking=True)
        opt = self.opt
        outputs = self.model.render(rays_o, rays_d, direction_norms, staged=False, bg_color=None, perturb=True,
                                    **{k: v for k, v in vars(opt).items() if k in ('dt_gamma', 'max_steps', 'force_all_rays')})
        pred_rgb = outputs['image']
        loss = opt.rgb_weight * self.criterion(pred_rgb, gt_rgb).mean()
        has_depth = (gt_depth > DEPTH_EPSILON).to(pred_rgb.dtype)
        depth_err = (torch.abs(outputs['depth'] - gt_depth) * has_depth).sum() / has_depth.sum().clamp(min=1)
        loss = loss + opt.depth_weight * depth_err
        if getattr(opt, 'feature_loss', False) and 'features' in data:
            gt_features = data['features'].to(dev, non_blocking=True)
            loss = loss + opt.feature_weight * F.l1_loss(outputs['semantic_features'][:, :gt_features.shape[1]], gt_features)
        has_sem = gt_semantic >= 0
        ce = F.cross_entropy(outputs['semantic'], gt_semantic.clamp(min=0), reduction='none')
        sem_loss = (ce * has_sem).sum() / has_sem.sum().clamp(min=1)
        loss = loss + opt.semantic_weight * sem_loss
        return pred_rgb, gt_rgb, loss

    def fused_step_available(self):
        """The fused step (march -> field -> composite -> loss kernel -> backward, no autograd graph) covers the
        stock configuration: marched renderer, MSE criterion, any number of classes the compositing supports
        (3 + C + F <= 1280 channels: the 606-class ScanNet label set with 512-d features included)."""
        return (self.fused_step and self.model.cuda_ray and isinstance(self.criterion, torch.nn.MSELoss)
                and 3 + self.model.semantic_classes + self.model.hidden_dim_semantic <= 1280 and self.model.training)

    def _fused_train_step(self, data):
        """train_step + backward of the reference (trainer.py:54-94) as six library calls: the losses and their
        gradients w.r.t. the composited outputs come from one kernel (al_loss_fwd_bwd), the rest is the fused
        renderer.  Numerically the same loss as train_step (tests/test_trainer_gpu.py)."""
        dev = self.device
        nb = dict(non_blocking=True)
        batch = {k: data[k].to(dev, **nb) for k in ('rays_o', 'rays_d', 'direction_norms', 'pixels', 'depth', 'semantic')}
        if getattr(self.opt, 'feature_loss', False) and 'features' in data:
            batch['features'] = data['features'].to(dev, **nb)
        kw = {k: v for k, v in vars(self.opt).items() if k in ('dt_gamma', 'max_steps', 'force_all_rays')}
        loss5, _ = self._fused_core(batch, kw, None, None)
        self.last_loss_parts = loss5
        return loss5[0]

    def _fused_core(self, batch, kw, counter, arena, capacity=None):
        """march -> field -> composite -> loss kernel -> backward on device tensors; `counter` (optional) replaces the
        model's rotating step counter row (graph capture: the row is copied back after the replay); `arena`
        (renderer.StepArena, optional) supplies every scratch buffer."""
        from . import renderer
        from ._lib import call, ptr, stream_ptr
        dev = self.device
        m, opt = self.model, self.opt
        rays_o, rays_d = batch['rays_o'], batch['rays_d']
        norms = batch['direction_norms'].reshape(-1).float().contiguous()
        gt_rgb = batch['pixels'].reshape(-1, 3).float().contiguous()
        gt_depth = batch['depth'].reshape(-1).float().contiguous()
        gt_sem = batch['semantic'].reshape(-1).long().contiguous()
        gt_feat = batch['features'].float().contiguous() if 'features' in batch else None
        c, rays_d = m.train_forward_raw(rays_o, rays_d, perturb=True, counter=counter, arena=arena, capacity=capacity, **kw)
        N, K = c.N, c.K
        A = arena if arena is not None else renderer._TorchAlloc(dev)
        loss5 = A.get('loss5', 5)
        counts = A.get('loss_counts', 2, torch.int32)
        g_ws, g_depth, g_out = A.get('g_ws', N), A.get('g_depth', N), A.get('g_out', (N, K))
        Fg = 0 if gt_feat is None else gt_feat.shape[1]
        call("al_loss_fwd_bwd", ptr(c.ws), ptr(c.depth), ptr(c.out), N, int(m.semantic_classes),
             int(m.hidden_dim_semantic), ptr(norms), ptr(gt_rgb), ptr(gt_depth), ptr(gt_sem), ptr(gt_feat), int(Fg),
             float(opt.rgb_weight), float(opt.depth_weight), float(opt.semantic_weight),
             float(getattr(opt, 'feature_weight', 0.0)), DEPTH_EPSILON, 1.0, ptr(loss5), ptr(counts), ptr(g_ws),
             ptr(g_depth), ptr(g_out), stream_ptr(dev))
        renderer.fused_train_backward(m, c, g_ws, g_depth, g_out, m.field_params(), arena)
        return loss5, m.last_meta

    # ------------------------------------------------------------ CUDA-graph replay of the fused step
    def _graph_train_step(self, data):
        """The fused step as ONE graph launch.  The step has no host synchronisation and fixed launch geometry (every
        kernel reads the live sample count from device memory), so march -> field -> composite -> loss -> backward is
        captured once per buffer CAPACITY and replayed on static buffers; the sample budget M (it changes whenever the
        occupancy refresh updates `mean_count`, every `update_interval` steps) is read by the march from device memory.
        The batch is copied (H2D or D2D) into static input tensors.  The optimiser step and the gradient all-reduce
        stay outside the graph."""
        dev = self.device
        m, opt = self.model, self.opt
        kw = {k: v for k, v in vars(opt).items() if k in ('dt_gamma', 'max_steps', 'force_all_rays')}
        max_steps = int(kw.get('max_steps', 1024))
        N = data['rays_o'].reshape(-1, 3).shape[0]
        use_feat = bool(getattr(opt, 'feature_loss', False) and 'features' in data)
        Fg = data['features'].shape[-1] if use_feat else 0
        M = m.sample_budget(N, max_steps, bool(kw.get('force_all_rays', False)))     # the reference's budget rule
        st = self._graph_state
        if st is None or st['shape'] != (N, Fg):
            from .renderer import StepArena
            f32 = dict(dtype=torch.float32, device=dev)
            packed = PackedBatch(N, Fg, dev)               # static inputs: views of one flat buffer
            st = {'shape': (N, Fg), 'key': None, 'graph': None, 'arena': StepArena(dev), 'cap': 0,
                  'in': dict(packed), 'in_packed': packed,
                  'counter': torch.zeros(2, dtype=torch.int32, device=dev)}
            self._graph_state = st
        if isinstance(data, PackedBatch) and data.shape_key == (N, Fg):
            st['in_packed'].flat.copy_(data.flat, non_blocking=True)      # ONE copy (H2D from pinned memory, or D2D)
        else:
            for k, buf in st['in'].items():
                buf.copy_(data[k].reshape(buf.shape), non_blocking=True)
        thresh = float(getattr(m, 'train_t_thresh', 0.0))
        if st.get('thresh') != thresh:
            st['thresh'], st['cap'], st['arena_cap'] = thresh, 0, 0   # other scratch buffers: one eager step on the arena first
        # The budget M moves with every occupancy refresh (mean of the last 16 steps' totals).  Buffers and launch
        # geometry are sized by a CAPACITY that only changes when M leaves [0.7 cap, cap]; M itself reaches the march
        # through device memory (al_march_rays_train_budget), so a refresh does not force a re-capture.
        if M > st['cap'] or M < 0.7 * st['cap']:
            new_cap = min(N * max_steps, -(-int(M * 1.1) // 32768) * 32768)
            new_cap = max(new_cap, M)
            grow = new_cap > st.get('arena_cap', 0)
            st['cap'], st['graph'], st['key'] = new_cap, None, None
            if grow:
                # the arena has to grow: run this step kernel by kernel on it (allocating), capture from the next step on
                st['arena_cap'] = new_cap
                m.budget_tensor(M)
                slot = m.local_step % 16
                m.local_step += 1
                st['counter'].zero_()
                loss5, meta = self._fused_core(dict(st['in']), kw, st['counter'], st['arena'], capacity=new_cap)
                m.step_counter[slot].copy_(st['counter'], non_blocking=True)
                self.last_loss_parts = loss5
                return loss5[0].clone()
        cap = st['cap']
        m.budget_tensor(M)                                 # outside the graph: the replay reads the fresh value
        # the optimiser joins the graph when it is the fused Adam without a gradient exchange in between (single GPU):
        # al_adam_multi reads step count and learning rate from device memory
        adam_in_graph = (self.grad_sync is None and len(self.optimizers) == 1 and isinstance(self.optimizer, FusedAdam)
                         and self.optimizer.zero_grad_in_step)
        from ._lib import lib as _l
        key = (N, Fg, cap, float(kw.get('dt_gamma', 0)), max_steps, thresh, bool(kw.get('force_all_rays', False)),
               int(_l.al_set_mlp_backend(-1)), adam_in_graph, id(self.optimizer))
        if st['key'] != key:
            st['graph'] = None
            g = torch.cuda.CUDAGraph()
            # capture on a side stream by hand: torch.cuda.graph() would also synchronise the device and empty the
            # allocator cache on every re-capture; nothing is allocated here (StepArena refuses to)
            side = st.setdefault('stream', torch.cuda.Stream(device=dev))
            side.wait_stream(torch.cuda.current_stream(dev))
            if adam_in_graph:
                self.optimizer._device_state()          # allocates the device scalars: not inside the capture
            n0 = _l.al_launch_count()
            with torch.cuda.stream(side):
                g.capture_begin()
                try:
                    st['counter'].zero_()
                    loss5, meta = self._fused_core(dict(st['in']), kw, st['counter'], st['arena'], capacity=cap)
                    if adam_in_graph:
                        self.optimizer.step_device(sync_lr=False)
                finally:
                    g.capture_end()
            torch.cuda.current_stream(dev).wait_stream(side)
            # kernels of this library recorded in the graph (al_launch_count counts enqueues, also while capturing)
            st.update(graph=g, key=key, loss5=loss5, meta=meta, kernels=int(_l.al_launch_count() - n0), adam=adam_in_graph)
            self.graph_kernel_launches -= st['kernels']    # recorded, not executed
            if adam_in_graph:                              # the capture advanced the host mirror of the step count
                for d in self.optimizer._device_state():
                    for p_ in d['params']:
                        self.optimizer.state[p_]['step'] -= 1
        slot = m.local_step % 16
        m.local_step += 1
        if st['adam']:
            self.optimizer.sync_device_lr()
        st['graph'].replay()
        if st['adam']:
            for d in self.optimizer._device_state():
                for p_ in d['params']:
                    self.optimizer.state[p_]['step'] += 1
            self._stepped_in_graph = True
        self.graph_kernel_launches += st['kernels']        # kernels of this library executed by graph replays
        m.step_counter[slot].copy_(st['counter'], non_blocking=True)
        m.last_meta = st['meta']
        self.last_loss_parts = st['loss5']
        return st['loss5'][0]

    def train_one_step(self, data):
        """zero_grad -> train_step -> backward -> (gradient all-reduce) -> optimiser step; occupancy refresh
        every `update_interval` steps.  Returns the (device) loss."""
        self._stepped_in_graph = False
        if self.model.cuda_ray and self.global_step % self.update_interval == 0:
            self.model.update_extra_state()
        for o in self.optimizers:
            o.zero_grad()
        if self.fused_step_available():
            # the first steps run eagerly (lazy kernel attributes, gradient buffers); then one graph launch per step
            if self.use_graph and self.global_step >= 2 and all(
                    p.grad is not None for p in self.model.field_params() if p is not None and p.requires_grad):
                loss = self._graph_train_step(data)
            else:
                loss = self._fused_train_step(data)
        else:
            _, _, loss = self.train_step(data)
            loss.backward()
        if self.grad_sync is not None:
            self.grad_sync()
        if not self._stepped_in_graph:
            for o in self.optimizers:
                o.step()
        self.global_step += 1
        self.last_loss = loss.detach()
        return self.last_loss

    def train_iterations(self, dataloader, iterations):
        self.model.train()
        data_src = getattr(dataloader, '_data', dataloader)
        if self.model.cuda_ray and hasattr(data_src, 'poses'):
            self.model.mark_untrained_grid(data_src.poses, data_src.intrinsics)
        iterator = iter(dataloader)
        for it in range(iterations):
            self.train_one_step(next(iterator))
            if self.log_interval and (it + 1) % self.log_interval == 0 and self.local_rank == 0:
                print(f"[{self.name}] step {self.global_step} loss {self.last_loss.item():.4f}", flush=True)
        if self.ema is not None:
            self.ema.update()
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()

    def train(self, dataloader, epochs, iterations_per_epoch=1000):
        for _ in range(epochs):
            self.train_iterations(dataloader, iterations_per_epoch)
            self.epoch += 1
            if self.workspace is not None:
                self.save_checkpoint()        # every rank calls (sharded optimiser state is gathered), rank 0 writes

    @torch.no_grad()
    def test_step(self, data):
        H, W = data['H'], data['W']
        outputs = self.model.render(data['rays_o'], data['rays_d'], data['direction_norms'], staged=True, perturb=False)
        pred_rgb = outputs['image'].reshape(-1, H, W, 3)
        pred_depth = outputs['depth'].reshape(-1, H, W)
        pred_semantic = outputs['semantic'].reshape(-1, H, W, outputs['semantic'].shape[-1])
        return pred_rgb, pred_depth, pred_semantic, outputs['semantic_features']

    @torch.no_grad()
    def eval_step(self, data):
        dev = self.device
        gt_rgb = data['pixels'].to(dev)
        H, W, _ = gt_rgb.shape
        outputs = self.model.render(data['rays_o'].to(dev), data['rays_d'].to(dev), data['direction_norms'].to(dev),
                                    staged=True, bg_color=None, perturb=False)
        pred_rgb = outputs['image'].reshape(H, W, 3)
        pred_depth = outputs['depth'].reshape(H, W)
        loss = self.criterion(pred_rgb, gt_rgb).mean()
        return pred_rgb, pred_depth, outputs['semantic'].reshape(H, W, -1), gt_rgb, loss

    # ------------------------------------------------------------ checkpoints (torch_ngp/nerf/utils.py:1124-1260)
    def save_checkpoint(self, name=None, full=True, remove_old=True):
        """The reference Trainer's checkpoint dictionary (utils.py:1133-1163): epoch, global_step, stats, precision,
        mean_count / mean_density (marched renderer), and with `full`: opt{i}, lr_sched{i}, scaler, ema; then model.
        COLLECTIVE when the optimiser is sharded over ranks (its state_dict() gathers the moments): every rank calls,
        rank 0 writes."""
        name = name or f'{self.name}_ep{self.epoch:04d}'
        state = {'epoch': self.epoch, 'global_step': self.global_step, 'stats': self.stats,
                 'precision': "half" if self.fp16 else "full"}
        if self.model.cuda_ray:
            state['mean_count'] = self.model.mean_count
            state['mean_density'] = self.model.mean_density
        if full:
            for i, o in enumerate(self.optimizers):
                state[f'opt{i}'] = o.state_dict()
            for i, sch in enumerate(self.lr_schedulers):
                state[f'lr_sched{i}'] = sch.state_dict()
            # no loss scaling on this path (fp32 gradients out of the kernels): a disabled GradScaler's state
            state['scaler'] = {}
            if self.ema is not None:
                state['ema'] = self.ema.state_dict()
        state['model'] = self.model.state_dict()
        if self.local_rank != 0 or self.workspace is None:
            return None
        path = os.path.join(self.workspace, 'checkpoints', f'{name}.pth')
        if remove_old:
            self.stats["checkpoints"].append(path)
            while len(self.stats["checkpoints"]) > self.max_keep_ckpt:
                old = self.stats["checkpoints"].pop(0)
                if os.path.exists(old):
                    os.remove(old)
        torch.save(state, path)
        return path

    def load_checkpoint(self, checkpoint=None, model_only=False):
        """utils.py:1196-1260.  Also reads the round-1 layout of this trainer ('optimizer' / 'lr_scheduler' keys)."""
        if checkpoint is None:
            ckpts = sorted(glob.glob(os.path.join(self.workspace, 'checkpoints', f'{self.name}_ep*.pth')))
            if not ckpts:
                return False
            checkpoint = ckpts[-1]
        state = torch.load(checkpoint, map_location=self.device, weights_only=False)
        if 'model' not in state:
            self.model.load_state_dict(state)
            return True
        self.model.load_state_dict(state['model'], strict=False)
        if self.ema is not None and 'ema' in state:
            self.ema.load_state_dict(state['ema'])
        if self.model.cuda_ray:
            if 'mean_count' in state:
                self.model.mean_count = state['mean_count']
            if 'mean_density' in state:
                self.model.mean_density = state['mean_density']
        if model_only:
            return True
        self.stats = state.get('stats', self.stats)
        self.epoch, self.global_step = state.get('epoch', 0), state.get('global_step', 0)
        for i, o in enumerate(self.optimizers):
            sd = state.get(f'opt{i}', state.get('optimizer') if i == 0 else None)
            if sd is not None:
                o.load_state_dict(sd)
        for i, sch in enumerate(self.lr_schedulers):
            sd = state.get(f'lr_sched{i}', state.get('lr_scheduler') if i == 0 else None)
            if sd is not None:
                sch.load_state_dict(sd)
                for o in self.optimizers:
                    for g, lr in zip(o.param_groups, sch.get_last_lr()):
                        g['lr'] = lr
        self._graph_state = None
        return True
