"""Fused optimizer step: `torch.optim.Adam` as SCGaussian constructs and steps it
(reference scene/gaussian_model.py:477 and :502, :512: `torch.optim.Adam(l, lr=0.0, eps=1e-15)` with one
parameter per named group; reference train.py:204-208: `optimizer.step()` then `optimizer_bg.step()`), with every
parameter group updated by ONE kernel launch (include/scgr.h: scgr_adam_step) instead of torch's multi-tensor
chain of ~12 launches per optimizer.

    from scgaussian_b200.optim import Adam, step_all
    gaussians.optimizer = Adam(l, lr=0.0, eps=1e-15)          # same arguments as torch.optim.Adam
    step_all(gaussians.optimizer, gaussians.optimizer_bg)      # both optimizers of the reference in one launch

`Adam` is a `torch.optim.Optimizer`: `param_groups` (lr scheduling, reference :514-527), `state` with torch's own
keys `step` / `exp_avg` / `exp_avg_sq` (what the reference's densification edits in place, :758-843), `state_dict`
/ `load_state_dict` (checkpoints, :83, :103) and `zero_grad` are inherited and interchangeable with torch's.
No CPU / eager fallback: the update runs in libscgr.so or raises.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict

import torch

from . import _lib
from ._lib import ADAM_MAX_GROUPS, ScgrAdamGroup, ScgrError, check


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *,
                 maximize=False, foreach=None, capturable=False, differentiable=False, fused=None):
        if weight_decay != 0 or amsgrad or maximize or capturable or differentiable:
            raise ScgrError("fused Adam implements what the reference uses: no weight_decay / amsgrad / maximize / "
                            "capturable / differentiable")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)

    def _collect(self, work, bumped):
        """Appends this optimizer's updates to `work`: (device, beta1, beta2, eps) -> [ScgrAdamGroup ...].
        Two passes: every parameter is validated (and its state created) first; only then are the step counts
        advanced -- recorded in `bumped` so that the caller can roll them back if the launch fails."""
        todo = []
        for group in self.param_groups:
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise ScgrError("fused Adam: weight_decay / amsgrad / maximize are not implemented")
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.device.type != "cuda":
                    raise ScgrError("fused Adam runs on CUDA tensors only (no CPU path exists)")
                if p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise ScgrError("fused Adam expects dense contiguous fp32 parameters")
                if p.numel() == 0:
                    continue
                grad = p.grad
                if grad.dtype != torch.float32 or not grad.is_contiguous():
                    grad = grad.float().contiguous()
                state = self.state[p]
                if len(state) == 0:     # torch/optim/adam.py _init_group
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if m.shape != p.shape or v.shape != p.shape or not m.is_contiguous() or not v.is_contiguous() \
                        or m.dtype != torch.float32 or v.dtype != torch.float32 or m.device != p.device:
                    raise ScgrError("fused Adam: optimizer state does not match its parameter (shape / dtype / device)")
                todo.append((p, grad, state, m, v, group, float(beta1), float(beta2)))
        for p, grad, state, m, v, group, beta1, beta2 in todo:
            state["step"] += 1
            bumped.append(state)
            step = int(state["step"].item()) if isinstance(state["step"], torch.Tensor) else int(state["step"])
            key = (p.device, beta1, beta2, float(group["eps"]))
            work[key].append((ScgrAdamGroup(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(),
                                            p.numel(), float(group["lr"]), step), grad))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        work, bumped = defaultdict(list), []
        try:
            self._collect(work, bumped)
            _launch(work)
        except Exception:
            _rollback(bumped)
            raise
        return loss


def _rollback(bumped) -> None:
    """A failed step leaves every step count where it was (no update was applied for it)."""
    for state in bumped:
        state["step"] -= 1


def _launch(work) -> None:
    lib = _lib.load()
    for (device, beta1, beta2, eps), items in work.items():
        with torch.cuda.device(device):
            stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            for i in range(0, len(items), ADAM_MAX_GROUPS):
                chunk = items[i:i + ADAM_MAX_GROUPS]
                table = (ScgrAdamGroup * len(chunk))(*[g for g, _ in chunk])
                check(lib.scgr_adam_step(table, len(chunk), beta1, beta2, eps, stream))


@torch.no_grad()
def step_all(*optimizers: Adam) -> None:
    """Steps several fused optimizers in one launch (reference train.py:204-208 steps `optimizer` and
    `optimizer_bg` back to back: 12 parameter groups, all with the same betas / eps)."""
    work, bumped = defaultdict(list), []
    for opt in optimizers:
        if not isinstance(opt, Adam):
            raise ScgrError("step_all takes scgaussian_b200.optim.Adam optimizers")
    try:
        for opt in optimizers:
            opt._collect(work, bumped)
        _launch(work)
    except Exception:
        _rollback(bumped)
        raise
