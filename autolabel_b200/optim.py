"""Fused Adam (al_adam_step) with torch.optim.Adam semantics — the optimiser configuration of the
reference's scripts/train.py:50-63 (lr 5e-3, betas (0.9, 0.99), eps 1e-15, L2 weight decay 1e-6 on
the MLP parameters only).  One kernel per parameter tensor: gradient unscale, moment update,
parameter update and gradient zeroing in a single pass (32 B/param of HBM traffic)."""
import torch

from ._lib import call, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, zero_grad_in_step=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.zero_grad_in_step = zero_grad_in_step
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
        return loss

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed inside step(); keep the buffers (they are accumulated into in place)."""
        if self.zero_grad_in_step:
            return
        super().zero_grad(set_to_none=set_to_none)
