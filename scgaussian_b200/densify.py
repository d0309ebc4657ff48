"""Prune compaction (second half): `GaussianModel.prune_points(mask)` of reference
scene/gaussian_model.py:795-820 with its helper `_prune_optimizer` (:777-793).

The reference evaluates `t[valid_points_mask]` once per per-Gaussian array -- 12 parameters, their `exp_avg` /
`exp_avg_sq`, `_rayo`, `_rayd`, `xyz_gradient_accum`, `denom`, `max_radii2D`: ~45 boolean-mask gathers, each a
nonzero + host synchronisation + gather kernel.  Here the two masks (ray-based set, free set) are turned into index
lists once, and every array that shares an index list is compacted by ONE launch (include/scgr.h: scgr_gather_rows).

    from scgaussian_b200.densify import prune_points
    prune_points(gaussians, prune_mask)          # instead of gaussians.prune_points(prune_mask)

Same effect on the model: new `nn.Parameter`s in `optimizer.param_groups` (one parameter per named group), optimizer
state re-keyed to them with compacted moments and the step count kept, model attributes replaced.  Works on
`torch.optim.Adam` and on `scgaussian_b200.optim.Adam` alike.  No CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import GATHER_MAX_ARRAYS, ScgrError, ScgrRowGather, check

# optimizer group name -> model attribute (reference scene/gaussian_model.py:493-510 and :803-817)
GROUP_ATTR = {"zval": "_zval", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
              "scaling": "_scaling", "rotation": "_rotation",
              "bg_xyz": "bg_xyz", "bg_f_dc": "bg_features_dc", "bg_f_rest": "bg_features_rest",
              "bg_opacity": "bg_opacity", "bg_scaling": "bg_scaling", "bg_rotation": "bg_rotation"}


def _require_cuda(device, what: str) -> None:
    if device.type != "cuda":
        raise ScgrError(f"{what} runs on CUDA tensors only (no CPU path exists)")


def gather_rows(tensors: List[torch.Tensor], index: torch.Tensor) -> List[torch.Tensor]:
    """[t[index] for t in tensors] for fp32 CUDA tensors that share dim 0, in one launch per 48 arrays.
    `index` is an int64 CUDA vector (e.g. `mask.nonzero().squeeze(1)`)."""
    if not tensors:
        return []
    lib = _lib.load()
    dev = index.device
    _require_cuda(dev, "gather_rows")
    if index.dtype != torch.int64 or index.dim() != 1:
        raise ScgrError("gather_rows: index must be a 1-D int64 tensor")
    index = index.contiguous()
    n_in, n_out = int(tensors[0].shape[0]), int(index.shape[0])
    srcs, outs, table = [], [], []
    for t in tensors:
        if t.device != dev or t.dtype != torch.float32 or t.dim() < 1 or t.shape[0] != n_in:
            raise ScgrError("gather_rows: tensors must be fp32, on the index's device, with a common first dimension")
        src = t.detach().contiguous()
        dst = torch.empty((n_out,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
        row = src.numel() // n_in if n_in else 0
        srcs.append(src)
        outs.append(dst)
        if row and n_out:
            table.append(ScgrRowGather(src.data_ptr(), dst.data_ptr(), row))
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for i in range(0, len(table), GATHER_MAX_ARRAYS):
            chunk = table[i:i + GATHER_MAX_ARRAYS]
            check(lib.scgr_gather_rows((ScgrRowGather * len(chunk))(*chunk), len(chunk), index.data_ptr(), n_out, stream))
    return outs


def prune_optimizer(index: torch.Tensor, optimizer, extra: Optional[List[torch.Tensor]] = None
                    ) -> Tuple[Dict[str, nn.Parameter], List[torch.Tensor]]:
    """reference scene/gaussian_model.py:777-793 `_prune_optimizer(mask, optimizer)` for `index = mask.nonzero()`:
    every group's parameter and moments compacted in one launch together with the `extra` tensors that share the
    mask.  Returns ({group name: new parameter}, compacted extras)."""
    extra = list(extra or [])
    jobs = []          # (group, stored_state or None)
    tensors = []
    for group in optimizer.param_groups:
        if len(group["params"]) != 1:
            raise ScgrError("prune_optimizer expects one parameter per group (reference scene/gaussian_model.py:824)")
        p = group["params"][0]
        st = optimizer.state.get(p, None)
        jobs.append((group, st))
        tensors.append(p)
        if st is not None:
            tensors += [st["exp_avg"], st["exp_avg_sq"]]
    outs = gather_rows(tensors + extra, index)
    optimizable = {}
    k = 0
    for group, st in jobs:
        old = group["params"][0]
        new = nn.Parameter(outs[k].requires_grad_(True))
        k += 1
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = outs[k], outs[k + 1]
            k += 2
            del optimizer.state[old]
            optimizer.state[new] = st
        group["params"][0] = new
        optimizable[group["name"]] = new
    return optimizable, outs[k:]


def prune_points(pc, mask: torch.Tensor) -> None:
    """reference scene/gaussian_model.py:795-820: drops the Gaussians where `mask` is True from the ray-based set
    (`pc.optimizer`, `_rayo`, `_rayd`), the free set (`pc.optimizer_bg`) and the densification statistics."""
    _require_cuda(mask.device, "prune_points")
    valid = ~mask.reshape(-1).to(torch.bool)
    n_ray = int(pc._zval.shape[0])
    idx_ray = valid[:n_ray].nonzero().squeeze(1)
    idx_bg = valid[n_ray:].nonzero().squeeze(1)
    tensors, (rayo, rayd) = prune_optimizer(idx_ray, pc.optimizer, [pc._rayo, pc._rayd])
    tensors_bg, _ = prune_optimizer(idx_bg, pc.optimizer_bg)
    pc._rayo, pc._rayd = rayo, rayd
    for name, t in list(tensors.items()) + list(tensors_bg.items()):
        if name not in GROUP_ATTR:
            raise ScgrError(f"prune_points: unknown optimizer group {name!r}")
        setattr(pc, GROUP_ATTR[name], t)
    pc.xyz_gradient_accum, pc.denom, pc.max_radii2D = gather_rows(
        [pc.xyz_gradient_accum, pc.denom, pc.max_radii2D], torch.cat([idx_ray, idx_bg + n_ray]))
