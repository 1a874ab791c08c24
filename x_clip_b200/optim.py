"""Fused AdamW for the CLIP module (the step after the hot path).

The reference ships no optimizer or training loop (README.md:44-58 calls `loss.backward()` and
stops); a user would put torch.optim.AdamW behind it - hundreds of small elementwise launches for
the ~150 parameter tensors of cfg3.  `FusedAdamW` moves every fp32 parameter of the module into ONE
flat buffer (the parameters become views, `state_dict` is unchanged) and runs a single
`xclip_adamw_step` launch over (param, grad, exp_avg, exp_avg_sq) per step; with a
`distributed.GradSync` the averaged bucket gradients are consumed where they are.
The update rule is torch.optim.AdamW's (decoupled weight decay, bias correction)."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import kernels as K


class FusedAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdamW: no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev or not p.is_cuda:
                raise ValueError("FusedAdamW: parameters must be fp32 CUDA tensors on one device")
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        # every parameter starts at a multiple of 4 elements (16 bytes) inside the flat buffers
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view                       # the module now trains inside the flat buffer
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.grad = torch.zeros_like(self.flat)

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            p.grad = None if set_to_none else (p.grad.zero_() if p.grad is not None else None)

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        """One AdamW step from the parameters' `.grad` (parameters without a gradient keep a zero
        gradient, i.e. they only see weight decay and moment decay - as with a zero grad in torch)."""
        self.step_count += 1
        self.grad.zero_()
        src = [p.grad.reshape(-1) for p in self.params if p.grad is not None]
        dst = [self.grad[off:off + p.numel()] for p, off in zip(self.params, self.offsets) if p.grad is not None]
        if src:
            torch._foreach_copy_(dst, src)
        K.adamw_step_(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0],
                      self.betas[1], self.eps, self.weight_decay, self.step_count, grad_scale)
