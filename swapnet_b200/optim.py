"""Fused AdamW over flat parameter / gradient / moment buffers.

Same update rule and hyper-parameters as the `torch.optim.AdamW(params, lr, weight_decay, betas)` the
reference builds (optimizers/__init__.py:48-59), one kernel launch per network instead of torch's
foreach pass (7 reads/writes of 140 M parameters).  `state_dict()` / `load_state_dict()` keep
torch.optim.AdamW's layout ('step', 'exp_avg', 'exp_avg_sq' per parameter), so `{epoch}_optim_{G,D}.pth`
files are interchangeable with the reference's (base_model.py:168-173,203-212).
"""
from __future__ import annotations

from typing import List

import torch

from . import ops


def flatten_parameters(params: List[torch.nn.Parameter]) -> torch.Tensor:
    """Move the parameters into ONE contiguous fp32 buffer; each p.data becomes a view of it (values,
    state_dict keys and in-place load_state_dict are unaffected)."""
    total = sum(p.numel() for p in params)
    flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view_as(p)
        off += n
    return flat


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, flat_param: torch.Tensor, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        assert flat_param.numel() == sum(p.numel() for p in params)
        off = 0
        for p in params:   # the parameters must be consecutive views of flat_param
            assert p.data_ptr() == flat_param.data_ptr() + 4 * off, "parameters are not views of the flat buffer"
            off += p.numel()
        self.flat_param = flat_param
        self.flat_grad = None                      # attached by the engine (Engine.alloc_grads)
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self._step = 0
        self._expose_state()

    def _expose_state(self) -> None:
        off = 0
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(self._step)),
                             "exp_avg": self.exp_avg[off:off + n].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p)}
            off += n

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None and self.flat_grad is not None, "FusedAdamW needs the engine's flat gradient buffer"
        g = self.param_groups[0]
        self._step += 1
        ops.adamw_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0],
                       g["betas"][1], g["eps"], g["weight_decay"], self._step)
        for st in self.state.values():
            st["step"].fill_(float(self._step))

    def advance(self, gscale: float = 1.0):
        """Host half of a step whose kernel reads its scalars from device memory (BaseGAN's step-parameter buffer):
        bump the step counter and return the 8 scalars of ops.adamw_step_dev for it."""
        g = self.param_groups[0]
        self._step += 1
        for st in self.state.values():
            st["step"].fill_(float(self._step))
        return ops.adamw_hyper(g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._step, gscale)

    @torch.no_grad()
    def launch(self, hyper_dev: torch.Tensor) -> None:
        """Device half: one fused kernel over the flat buffers (capturable: no host state is touched)."""
        assert self.flat_grad is not None, "FusedAdamW needs the engine's flat gradient buffer"
        ops.adamw_step_dev(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, hyper_dev)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # torch replaced the per-parameter tensors: copy them back into the flat buffers
        off = 0
        step = 0
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            st = self.state.get(p, {})
            if "exp_avg" in st:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = int(float(st["step"]))
            off += n
        self._step = step
        self._expose_state()
