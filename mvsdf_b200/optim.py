"""Fused optimiser step of the training loop: torch.optim.Adam.step
(code/training/idr_train.py:113, :300; Adam defaults, no weight decay) with the reference's gradient-norm report and
torch.nn.utils.clip_grad_norm_ (:289-294) folded into the same two launches (mvsdf_adam_step: one pass for the squared
norm over all tensors, one for the update)."""
from __future__ import annotations

from ctypes import c_int64, c_void_p
from typing import Iterable, Optional

import torch

from . import _lib


class B200Adam(torch.optim.Optimizer):
    """Drop-in for ``torch.optim.Adam(model.parameters(), lr=...)`` on CUDA fp32 parameters.

    ``step(max_grad_norm=None)``: with a value, gradients are scaled by min(1, max_norm / (||g|| + 1e-6)) exactly like
    ``clip_grad_norm_`` before they enter the moments; ``grad_norm`` (device scalar) holds ||g|| of the last step, which the
    reference prints every iteration (idr_train.py:289-290) -- read it only when needed, it costs a host sync.  Parameters
    whose ``.grad`` is None are skipped, like torch does."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.grad_norm: Optional[torch.Tensor] = None
        self._scratch: Optional[torch.Tensor] = None

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        assert closure is None, "closures are not supported"
        L = _lib.lib()
        ps, gs, ms, vs, sizes = [], [], [], [], []
        step_no, hyper, dev = None, None, None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.MvsdfError("B200Adam: parameters must be contiguous CUDA fp32 tensors (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                this = (group["lr"], group["betas"][0], group["betas"][1], group["eps"])
                if hyper is None:
                    hyper, step_no, dev = this, st["step"] + 1, p.device
                elif this != hyper or st["step"] + 1 != step_no:
                    raise _lib.MvsdfError("B200Adam: one hyper-parameter set / step count for all parameters (the reference uses a "
                                          "single param group and every parameter receives a gradient, idr_train.py:113)")
                ps.append(p)
                gs.append(g)
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
                sizes.append(p.numel())
        if not ps:
            return None
        for p in ps:
            self.state[p]["step"] = step_no
        if self._scratch is None or self._scratch.device != dev:
            self._scratch = torch.zeros(2, dtype=torch.float64, device=dev)
            self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
        arr = (c_int64 * len(sizes))(*sizes)
        self._keep = gs
        _lib.check(L.mvsdf_adam_step(len(ps), _lib.ptr_array(ps), _lib.ptr_array(gs), _lib.ptr_array(ms), _lib.ptr_array(vs), arr,
                                     float(hyper[0]), float(hyper[1]), float(hyper[2]), float(hyper[3]), int(step_no),
                                     float(max_grad_norm) if max_grad_norm else 0.0, _lib.ptr(self._scratch), _lib.ptr(self.grad_norm),
                                     c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return None
