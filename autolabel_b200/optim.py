"""Fused Adam (al_adam_step / al_adam_multi) with torch.optim.Adam semantics — the optimiser configuration of the
reference's scripts/train.py:50-63 (lr 5e-3, betas (0.9, 0.99), eps 1e-15, L2 weight decay 1e-6 on
the MLP parameters only).  Gradient unscale, moment update, parameter update and gradient zeroing in a
single pass (32 B/param of HBM traffic).  `step()` launches one kernel per parameter tensor with the step count
on the host; `step_device()` is ONE launch for all tensors with the step count and the learning rate in device
memory — the form the graph-captured training step replays (tests/test_optim_gpu.py pins both on torch.optim.Adam)."""
import ctypes

import torch

from ._lib import AdamTensor, call, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, zero_grad_in_step=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.zero_grad_in_step = zero_grad_in_step
        self._dev_state = None
        self.grad_scale = 1.0  # multiplied into every gradient (1/loss_scale, 1/world_size, ...)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam needs CUDA parameters; there is no CPU fallback")
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['step'] += 1
                call("al_adam_step", ptr(p), ptr(p.grad), ptr(st['exp_avg']), ptr(st['exp_avg_sq']), p.numel(),
                     float(group['lr']), float(b1), float(b2), float(group['eps']), float(group['weight_decay']),
                     int(st['step']), float(self.grad_scale), 1 if self.zero_grad_in_step else 0,
                     stream_ptr(p.device))
        for d in (getattr(self, '_dev_state', None) or []):
            d['step'].add_(1)                       # keep the device-side counters of step_device() in step
        return loss

    # ------------------------------------------------------------ graph-resident form
    def _device_state(self):
        """Per (betas, eps) group of tensors: the al_adam_tensor_t array, the device step counter and learning rate."""
        ds = self._dev_state
        if ds is not None:
            return ds
        ds = []
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.requires_grad]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam needs CUDA parameters; there is no CPU fallback")
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            dev = ps[0].device
            for i in range(0, len(ps), 8):
                chunk = ps[i:i + 8]
                steps = {int(self.state[p]['step']) for p in chunk}
                if len(steps) != 1:
                    raise RuntimeError("step_device(): the tensors of a group must share one step count")
                arr = (AdamTensor * len(chunk))()
                for j, p in enumerate(chunk):
                    st = self.state[p]
                    arr[j] = AdamTensor(p.data_ptr(), p.grad.data_ptr(), st['exp_avg'].data_ptr(),
                                        st['exp_avg_sq'].data_ptr(), p.numel(), float(group['weight_decay']))
                ds.append({'group': group, 'params': chunk, 'arr': arr, 'ptrs': [(p.data_ptr(), p.grad.data_ptr()) for p in chunk],
                           'step': torch.full((1,), steps.pop(), dtype=torch.int32, device=dev),
                           'lr': torch.full((1,), float(group['lr']), dtype=torch.float32, device=dev),
                           'lr_host': float(group['lr'])})
        self._dev_state = ds
        return ds

    def sync_device_lr(self):
        """Push a changed learning rate (lr scheduler) to the device scalars; call OUTSIDE graph capture."""
        for d in self._device_state():
            lr = float(d['group']['lr'])
            if lr != d['lr_host']:
                d['lr'].fill_(lr)
                d['lr_host'] = lr

    @torch.no_grad()
    def step_device(self, sync_lr=True):
        """One al_adam_multi launch per group of <= 8 tensors; nothing but device memory is read, so the call can be
        captured into a CUDA graph (pass sync_lr=False while capturing and call sync_device_lr() before each replay)."""
        ds = self._device_state()
        if sync_lr:
            self.sync_device_lr()
        for d in ds:
            if any((p.data_ptr(), p.grad.data_ptr()) != q for p, q in zip(d['params'], d['ptrs'])):
                raise RuntimeError("step_device(): a parameter or gradient buffer moved; call reset_device_state()")
            b1, b2 = d['group']['betas']
            call("al_adam_multi", d['arr'], len(d['params']), ptr(d['lr']), ptr(d['step']), float(b1), float(b2),
                 float(d['group']['eps']), float(self.grad_scale), 1 if self.zero_grad_in_step else 0,
                 stream_ptr(d['params'][0].device))
            for p in d['params']:
                self.state[p]['step'] += 1          # host mirror (state_dict, step())
        return None

    def reset_device_state(self):
        self._dev_state = None

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        self._dev_state = None

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed inside step(); keep the buffers (they are accumulated into in place)."""
        if self.zero_grad_in_step:
            return
        super().zero_grad(set_to_none=set_to_none)
