"""Adam as the reference implements it (minitorch/optim.py:30-76) on device-resident parameters.

Faithful to the reference including its quirk: the second moment is updated with (1 - beta1), not (1 - beta2)
(optim.py:68) -- kept so that loss curves are comparable."""
from __future__ import annotations

import math
from typing import Sequence

import torch


class Adam:
    """`flat=(values, grads)` (Module.flatten_parameters) switches to one fused kernel launch over the whole model
    (fa_b200_adam_step) instead of ~12 elementwise launches per parameter."""

    def __init__(self, parameters: Sequence, lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                 flat=None):
        self.parameters = list(parameters)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self._states = {id(p): {} for p in self.parameters}
        self.flat = flat
        if flat is not None:
            self._m, self._v, self._t = torch.zeros_like(flat[0]), torch.zeros_like(flat[0]), 0

    def zero_grad(self) -> None:
        if self.flat is not None:
            self.flat[1].zero_()  # the .grad views stay attached
            return
        for p in self.parameters:
            p.value.grad = None

    @torch.no_grad()
    def _flat_step(self) -> None:
        p, g = self.flat
        self._t += 1
        step_size = self.lr * math.sqrt(1.0 - self.beta2 ** self._t) / (1.0 - self.beta1 ** self._t)
        if p.is_cuda and p.dtype == torch.float32:
            from .. import _lib
            _lib.check(_lib.load_library().fa_b200_adam_step(p.data_ptr(), g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                                             p.numel(), step_size, self.beta1, self.beta2, 1.0 - self.beta1,
                                                             self.eps, torch.cuda.current_stream().cuda_stream))
            return
        self._m.mul_(self.beta1).add_(g, alpha=1 - self.beta1)
        self._v.mul_(self.beta2).addcmul_(g, g, value=1 - self.beta1)  # sic, optim.py:68
        p.addcdiv_(self._m, self._v.sqrt().add_(self.eps), value=-step_size)

    @torch.no_grad()
    def step(self) -> None:
        if self.flat is not None:
            return self._flat_step()
        for p in self.parameters:
            grad = p.value.grad
            if grad is None:
                continue
            state = self._states[id(p)]
            if not state:
                state["step"] = 0
                state["exp_avg"] = torch.zeros_like(grad)
                state["exp_avg_sq"] = torch.zeros_like(grad)
            state["step"] += 1
            state["exp_avg"].mul_(self.beta1).add_(grad, alpha=1 - self.beta1)
            state["exp_avg_sq"].mul_(self.beta2).addcmul_(grad, grad, value=1 - self.beta1)  # sic: (1 - beta1), optim.py:68
            denom = state["exp_avg_sq"].sqrt().add_(self.eps)
            bc1 = 1.0 - self.beta1 ** state["step"]
            bc2 = 1.0 - self.beta2 ** state["step"]
            step_size = self.lr * math.sqrt(bc2) / bc1
            p.value.addcdiv_(state["exp_avg"], denom, value=-step_size)
