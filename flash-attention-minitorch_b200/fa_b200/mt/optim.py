"""Adam as the reference implements it (minitorch/optim.py:30-76) on device-resident parameters.

Faithful to the reference including its quirk: the second moment is updated with (1 - beta1), not (1 - beta2)
(optim.py:68) -- kept so that loss curves are comparable."""
from __future__ import annotations

import math
from typing import Sequence

import torch


class Adam:
    def __init__(self, parameters: Sequence, lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
        self.parameters = list(parameters)
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self._states = {id(p): {} for p in self.parameters}

    def zero_grad(self) -> None:
        for p in self.parameters:
            p.value.grad = None

    @torch.no_grad()
    def step(self) -> None:
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
