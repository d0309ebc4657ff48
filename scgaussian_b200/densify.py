"""Densification bookkeeping: `GaussianModel.prune_points(mask)` of reference
scene/gaussian_model.py:795-820 with its helper `_prune_optimizer` (:777-793), `densification_postfix` /
`cat_tensors_to_optimizer` (:822-862) and the clone / split selection around them (`densify_and_clone`,
`densify_and_split`, `densify_and_prune`, :864-931).

The reference evaluates `t[valid_points_mask]` once per per-Gaussian array -- 12 parameters, their `exp_avg` /
`exp_avg_sq`, `_rayo`, `_rayd`, `xyz_gradient_accum`, `denom`, `max_radii2D`: ~45 boolean-mask gathers, each a
nonzero + host synchronisation + gather kernel.  Here the two masks (ray-based set, free set) are turned into index
lists once, and every array that shares an index list is compacted by ONE launch (include/scgr.h: scgr_gather_rows).

    from scgaussian_b200.densify import prune_points, densification_postfix
    prune_points(gaussians, prune_mask)          # instead of gaussians.prune_points(prune_mask)
    densification_postfix(gaussians, new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling,
                          new_rotation)          # instead of gaussians.densification_postfix(...): one launch

The append side (reference :822-862 `cat_tensors_to_optimizer` / `densification_postfix`: 18 torch.cat + 15 zero
fills) is one launch over flat segments (include/scgr.h: scgr_copy_segments).

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
from ._lib import COPY_MAX_SEGMENTS, GATHER_MAX_ARRAYS, ScgrError, ScgrRowGather, ScgrSegmentCopy, check

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


def _copy_segments(segments: List[Tuple[torch.Tensor, Optional[torch.Tensor]]], device) -> None:
    """dst.flatten()[:] = src.flatten() (src None: zeros) for every pair, in one launch per 48 segments."""
    lib = _lib.load()
    table = []
    for dst, src in segments:
        if not dst.is_contiguous() or dst.dtype != torch.float32 or (src is not None and (
                not src.is_contiguous() or src.dtype != torch.float32 or src.numel() != dst.numel())):
            raise ScgrError("copy_segments: fp32 contiguous tensors of equal size expected")
        if dst.numel():
            table.append(ScgrSegmentCopy(dst.data_ptr(), None if src is None else src.data_ptr(), dst.numel()))
    with torch.cuda.device(device):
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        for i in range(0, len(table), COPY_MAX_SEGMENTS):
            chunk = table[i:i + COPY_MAX_SEGMENTS]
            check(lib.scgr_copy_segments((ScgrSegmentCopy * len(chunk))(*chunk), len(chunk), stream))


def cat_tensors_to_optimizer(tensors_dict: Dict[str, torch.Tensor], optimizer,
                             zeros: Optional[List[torch.Tensor]] = None) -> Dict[str, nn.Parameter]:
    """reference scene/gaussian_model.py:822-842: every group's parameter extended with `tensors_dict[name]`, its
    moments with zeros; `zeros` are extra tensors to clear in the same launch."""
    device = optimizer.param_groups[0]["params"][0].device
    _require_cuda(device, "cat_tensors_to_optimizer")
    segments, jobs = [], []
    for group in optimizer.param_groups:
        if len(group["params"]) != 1:
            raise ScgrError("cat_tensors_to_optimizer expects one parameter per group (reference scene/gaussian_model.py:824)")
        old = group["params"][0]
        ext = tensors_dict[group["name"]]
        if ext.device != device or tuple(ext.shape[1:]) != tuple(old.shape[1:]):
            raise ScgrError(f"cat_tensors_to_optimizer: extension of group {group['name']!r} does not match its parameter")
        ext = ext.detach().to(torch.float32).contiguous()
        n_old, n_new = int(old.shape[0]), int(ext.shape[0])
        shape = (n_old + n_new,) + tuple(old.shape[1:])
        st = optimizer.state.get(old, None)

        def extended(src_old, src_new):
            t = torch.empty(shape, dtype=torch.float32, device=device)
            segments.append((t[:n_old], src_old.detach().contiguous()))
            segments.append((t[n_old:], src_new))
            return t
        new_p = extended(old, ext)
        moments = (extended(st["exp_avg"], None), extended(st["exp_avg_sq"], None)) if st is not None else None
        jobs.append((group, old, st, new_p, moments))
    for z in zeros or []:
        segments.append((z, None))
    _copy_segments(segments, device)
    optimizable = {}
    for group, old, st, new_p, moments in jobs:
        new = nn.Parameter(new_p.requires_grad_(True))
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = moments
            del optimizer.state[old]
            optimizer.state[new] = st
        group["params"][0] = new
        optimizable[group["name"]] = new
    return optimizable


def densification_postfix(pc, new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling,
                          new_rotation) -> None:
    """reference scene/gaussian_model.py:844-862: the new Gaussians join the free set (`pc.optimizer_bg`) and the
    densification statistics restart from zero."""
    d = {"bg_xyz": new_xyz, "bg_f_dc": new_features_dc, "bg_f_rest": new_features_rest, "bg_opacity": new_opacities,
         "bg_scaling": new_scaling, "bg_rotation": new_rotation}
    device = new_xyz.device
    _require_cuda(device, "densification_postfix")
    P = int(pc._zval.shape[0]) + int(pc.bg_xyz.shape[0]) + int(new_xyz.shape[0])
    accum = torch.empty(P, 1, dtype=torch.float32, device=device)
    denom = torch.empty(P, 1, dtype=torch.float32, device=device)
    max_radii = torch.empty(P, dtype=torch.float32, device=device)
    tensors = cat_tensors_to_optimizer(d, pc.optimizer_bg, zeros=[accum, denom, max_radii])
    for name, t in tensors.items():
        setattr(pc, GROUP_ATTR[name], t)
    pc.xyz_gradient_accum, pc.denom, pc.max_radii2D = accum, denom, max_radii


# ---------------------------------------------------------------------------------------------------------------
# Clone / split selection (rest): reference scene/gaussian_model.py:864-931, run every
# `densification_interval` iterations.  The reference gathers `torch.cat([set0, set1])[mask]` array by array (12 cats +
# 12 boolean-mask gathers per call, then the 18 cats + 15 zero fills of densification_postfix, then the ~45 gathers of
# prune_points).  Here the selection mask becomes two index lists once; each set's arrays are gathered by ONE launch
# (scgr_gather_rows), appended by ONE launch (scgr_copy_segments) and pruned by one launch per index list.  Thresholds,
# `torch.normal` sampling (same call, same shapes: same random stream as the reference) and the 3x3 rotation of the
# samples are host PyTorch on the selected rows only.
# ---------------------------------------------------------------------------------------------------------------
def _activated_scaling(pc) -> torch.Tensor:
    return torch.exp(torch.cat([pc._scaling.detach(), pc.bg_scaling.detach()]))      # reference :105-110


def _selected_rows(pc, mask: torch.Tensor):
    """(xyz, features_dc, features_rest, opacity, scaling, rotation) of the Gaussians where `mask` is set, ray-based set
    first -- `torch.cat([pc._x, pc.bg_x])[mask]` for every array of reference :904-909, two launches in all."""
    n_ray = int(pc._zval.shape[0])
    idx_ray = mask[:n_ray].nonzero().squeeze(1)
    idx_bg = mask[n_ray:].nonzero().squeeze(1)
    ray = gather_rows([pc._rayo, pc._rayd, pc._zval, pc._features_dc, pc._features_rest, pc._opacity, pc._scaling,
                       pc._rotation], idx_ray)
    bg = gather_rows([pc.bg_xyz, pc.bg_features_dc, pc.bg_features_rest, pc.bg_opacity, pc.bg_scaling, pc.bg_rotation], idx_bg)
    xyz_ray = ray[0] + ray[1] * ray[2]                               # reference :113-116 get_xyz of the ray-based set
    return [torch.cat([a, b]) for a, b in zip([xyz_ray] + ray[3:], bg)]


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """reference utils/general_utils.py:84-105 (the quaternion is normalised first), on the tensor's own device."""
    q = r / torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def replace_tensor_to_optimizer(tensor: torch.Tensor, name: str, optimizer) -> Dict[str, nn.Parameter]:
    """reference scene/gaussian_model.py:758-775: the named group's parameter replaced, its moments cleared."""
    out = {}
    for group in optimizer.param_groups:
        if group["name"] != name:
            continue
        old = group["params"][0]
        st = optimizer.state.get(old, None)
        new = nn.Parameter(tensor.requires_grad_(True))
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(tensor), torch.zeros_like(tensor)
            del optimizer.state[old]
            optimizer.state[new] = st
        group["params"][0] = new
        out[name] = new
    return out


def densify_and_clone(pc, grads: torch.Tensor, grad_threshold: float, scene_extent: float) -> None:
    """reference scene/gaussian_model.py:898-912: small Gaussians with a large view-space gradient are duplicated."""
    mask = (torch.norm(grads, dim=-1) >= grad_threshold) & \
           (torch.max(_activated_scaling(pc), dim=1).values <= pc.percent_dense * scene_extent)
    densification_postfix(pc, *_selected_rows(pc, mask))


def densify_and_split(pc, grads: torch.Tensor, grad_threshold: float, scene_extent: float, N: int = 2) -> None:
    """reference scene/gaussian_model.py:864-896: large Gaussians with a large gradient are replaced by N samples drawn
    from them (scale / (0.8 N)); a ray-based one stays in place, shrunk, instead of being pruned."""
    P = int(pc._zval.shape[0]) + int(pc.bg_xyz.shape[0])
    n_ray = int(pc._zval.shape[0])
    device = pc._zval.device
    padded = torch.zeros(P, dtype=grads.dtype, device=device)
    padded[:grads.shape[0]] = grads.squeeze()
    mask = (padded >= grad_threshold) & (torch.max(_activated_scaling(pc), dim=1).values > pc.percent_dense * scene_extent)
    xyz, f_dc, f_rest, opacity, scaling_raw, rotation = _selected_rows(pc, mask)
    scaling = torch.exp(scaling_raw)
    stds = scaling.repeat(N, 1)
    means = torch.zeros((stds.size(0), 3), dtype=stds.dtype, device=device)
    samples = torch.normal(mean=means, std=stds)                     # (same call as reference :875: same random stream)
    rots = build_rotation(rotation).repeat(N, 1, 1)
    new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + xyz.repeat(N, 1)
    new_scaling = torch.log(scaling.repeat(N, 1) / (0.8 * N))
    densification_postfix(pc, new_xyz, f_dc.repeat(N, 1, 1), f_rest.repeat(N, 1, 1), opacity.repeat(N, 1), new_scaling,
                          rotation.repeat(N, 1))
    n_split = int(mask.sum())
    shrunk = pc._scaling.detach().clone()
    shrunk[mask[:n_ray]] /= 0.8 * N
    pc._scaling = replace_tensor_to_optimizer(shrunk, "scaling", pc.optimizer)["scaling"]
    mask = mask.clone()
    mask[:n_ray] = False
    prune_points(pc, torch.cat((mask, torch.zeros(N * n_split, dtype=torch.bool, device=device))))


def densify_and_prune(pc, max_grad: float, min_opacity: float, extent: float, max_screen_size) -> None:
    """reference scene/gaussian_model.py:914-931 (`torch.cuda.empty_cache()` left to the caller)."""
    grads = pc.xyz_gradient_accum / pc.denom
    grads[grads.isnan()] = 0.0
    densify_and_clone(pc, grads, max_grad, extent)
    densify_and_split(pc, grads, max_grad, extent)
    n_ray = int(pc._zval.shape[0])
    opacity = torch.sigmoid(torch.cat([pc._opacity.detach(), pc.bg_opacity.detach()]))
    prune_mask = (opacity < min_opacity).squeeze(-1)
    if max_screen_size:
        big_vs = pc.max_radii2D > 1.5 * max_screen_size
        big_ws = _activated_scaling(pc).max(dim=1).values > 0.2 * extent
        prune_mask = prune_mask | big_vs | big_ws
    prune_mask[:n_ray] = False
    prune_points(pc, prune_mask)
