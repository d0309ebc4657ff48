"""Fused activations + hybrid assembly: what SCGaussian's
`GaussianModel.get_xyz / get_scaling / get_rotation / get_opacity / get_features` compute on every
render() call (reference scene/gaussian_model.py:105-152), as ONE kernel forward and ONE backward
(include/scgr.h: scgr_assemble_forward / scgr_assemble_backward) instead of ~20 torch kernels plus their
autograd nodes.  Opt-in: the reference's own properties keep working unchanged on top of the operator; a
trainer that wants the fused path calls

    means3D, scales, rotations, opacities, shs = assemble_model(gaussians)       # reads the raw parameters
    out = render(viewpoint_cam, gaussians, pipe, background)                     # the reference's render(), fused

`render` mirrors reference gaussian_renderer/__init__.py:20-118 (same arguments, same returned dict).
No CPU / eager fallback: every byte of compute goes through the C ABI of libscgr.so.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch

from . import _lib
from ._lib import (ScgrActivated, ScgrActivatedGrads, ScgrError, ScgrModel, ScgrModelGrads, ScgrModelSet,
                   ScgrModelSetGrads, check)

# argument order of _Assemble.apply: the ray-based set, then the free ("bg_") set
_RAY = ("rayo", "rayd", "zval", "scaling", "rotation", "opacity", "features_dc", "features_rest")
_BG = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
_TAIL = {"rayo": (3,), "rayd": (3,), "zval": (1,), "xyz": (3,), "scaling": (3,), "rotation": (4,), "opacity": (1,)}


def _prep(name: str, t: Optional[torch.Tensor], n: int, device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.device.type != "cuda":
        raise ScgrError("the fused model assembly runs on CUDA tensors only (no CPU path exists)")
    if t.device != device:
        raise ScgrError(f"{name} is on {t.device}, expected {device}")
    if t.shape[0] != n:
        raise ScgrError(f"{name} has {t.shape[0]} rows, expected {n}")
    if name in _TAIL and tuple(t.shape[1:]) != _TAIL[name]:
        raise ScgrError(f"{name} has shape {tuple(t.shape)}, expected [{n}, {_TAIL[name][0]}]")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t: Optional[torch.Tensor]):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


def _set(n: int, d: dict) -> ScgrModelSet:
    if n == 0:
        return ScgrModelSet(0, None, None, None, None, None, None, None, None, None)
    return ScgrModelSet(n, _p(d.get("xyz")), _p(d.get("rayo")), _p(d.get("rayd")), _p(d.get("zval")),
                        _p(d["scaling"]), _p(d["rotation"]), _p(d["opacity"]), _p(d["features_dc"]),
                        _p(d.get("features_rest")))


def _sh_rest(n: int, d: dict) -> Optional[int]:
    if n == 0:
        return None
    dc, rest = d["features_dc"], d["features_rest"]
    if dc.dim() != 3 or tuple(dc.shape[1:]) != (1, 3):
        raise ScgrError(f"features_dc has shape {tuple(dc.shape)}, expected [{n}, 1, 3]")
    if rest is None:
        return 0
    if rest.dim() != 3 or rest.shape[2] != 3:
        raise ScgrError(f"features_rest has shape {tuple(rest.shape)}, expected [{n}, K-1, 3]")
    return int(rest.shape[1])


class _Assemble(torch.autograd.Function):
    @staticmethod
    def forward(ctx, with_sh, *args):
        lib = _lib.load()
        ray_in = dict(zip(_RAY, args[:len(_RAY)]))
        bg_in = dict(zip(_BG, args[len(_RAY):]))
        n_ray = 0 if ray_in["scaling"] is None else int(ray_in["scaling"].shape[0])
        n_bg = 0 if bg_in["scaling"] is None else int(bg_in["scaling"].shape[0])
        if n_ray + n_bg == 0:
            raise ScgrError("assemble: the model holds no Gaussians")
        first = ray_in["scaling"] if n_ray else bg_in["scaling"]
        device = first.device
        ray = {k: _prep(k, v, n_ray, device) for k, v in ray_in.items()} if n_ray else {}
        bg = {k: _prep(k, v, n_bg, device) for k, v in bg_in.items()} if n_bg else {}
        for need, d, n in ((_RAY, ray, n_ray), (_BG, bg, n_bg)):
            missing = [k for k in need if n and d.get(k) is None and k != "features_rest"]
            if missing:
                raise ScgrError(f"assemble: missing raw parameters {missing}")
        rests = {r for r in (_sh_rest(n_ray, ray), _sh_rest(n_bg, bg)) if r is not None}
        if len(rests) != 1:
            raise ScgrError("assemble: the two sets disagree on the number of SH coefficients")
        sh_rest = rests.pop()
        P = n_ray + n_bg
        model = ScgrModel(sh_rest, (ScgrModelSet * 2)(_set(n_ray, ray), _set(n_bg, bg)))
        with torch.cuda.device(device):
            means3D = torch.empty(P, 3, dtype=torch.float32, device=device)
            scales = torch.empty(P, 3, dtype=torch.float32, device=device)
            rotations = torch.empty(P, 4, dtype=torch.float32, device=device)
            opacities = torch.empty(P, 1, dtype=torch.float32, device=device)
            # with_sh False: the operator reads features_dc / features_rest itself (split SH layout), no [P,K,3] copy
            shs = torch.empty(P, sh_rest + 1, 3, dtype=torch.float32, device=device) if with_sh else None
            out = ScgrActivated(means3D.data_ptr(), scales.data_ptr(), rotations.data_ptr(), opacities.data_ptr(),
                                _p(shs))
            stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            check(lib.scgr_assemble_forward(C.byref(model), C.byref(out), stream))
        ctx.sizes = (n_ray, n_bg, sh_rest)
        ctx.with_sh = bool(with_sh)
        ctx.layout = [k for k in _RAY if ray.get(k) is not None], [k for k in _BG if bg.get(k) is not None]
        ctx.save_for_backward(*[ray[k] for k in ctx.layout[0]], *[bg[k] for k in ctx.layout[1]])
        return means3D, scales, rotations, opacities, shs

    @staticmethod
    def backward(ctx, g_means3D, g_scales, g_rotations, g_opacities, g_shs):
        lib = _lib.load()
        n_ray, n_bg, sh_rest = ctx.sizes
        saved = list(ctx.saved_tensors)
        ray = dict(zip(ctx.layout[0], saved[:len(ctx.layout[0])]))
        bg = dict(zip(ctx.layout[1], saved[len(ctx.layout[0]):]))
        device = saved[0].device
        model = ScgrModel(sh_rest, (ScgrModelSet * 2)(_set(n_ray, ray), _set(n_bg, bg)))
        gin = [None if t is None else t.to(device=device, dtype=torch.float32).contiguous()
               for t in (g_means3D, g_scales, g_rotations, g_opacities, g_shs if ctx.with_sh else None)]
        if any(t is None for t in gin[:4]) or (ctx.with_sh and gin[4] is None):
            raise ScgrError("assemble backward: a gradient of the assembled arrays is missing")
        grads = ScgrActivatedGrads(*[_p(t) for t in gin])

        def new(n, *tail):
            return torch.empty(n, *tail, dtype=torch.float32, device=device) if n else None

        with torch.cuda.device(device):
            sh = ctx.with_sh      # split SH layout: the rasterizer's own node returns dL/dfeatures_*
            d_ray = {"zval": new(n_ray, 1), "scaling": new(n_ray, 3), "rotation": new(n_ray, 4),
                     "opacity": new(n_ray, 1), "features_dc": new(n_ray, 1, 3) if sh else None,
                     "features_rest": new(n_ray, sh_rest, 3) if sh else None}
            d_bg = {"xyz": new(n_bg, 3), "scaling": new(n_bg, 3), "rotation": new(n_bg, 4),
                    "opacity": new(n_bg, 1), "features_dc": new(n_bg, 1, 3) if sh else None,
                    "features_rest": new(n_bg, sh_rest, 3) if sh else None}

            def gset(d):
                return ScgrModelSetGrads(_p(d.get("xyz")), _p(d.get("zval")), _p(d["scaling"]), _p(d["rotation"]),
                                         _p(d["opacity"]), _p(d["features_dc"]), _p(d["features_rest"]))

            out = ScgrModelGrads((ScgrModelSetGrads * 2)(gset(d_ray), gset(d_bg)))
            stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            check(lib.scgr_assemble_backward(C.byref(model), C.byref(grads), C.byref(out), stream))
        # rayo / rayd are the fixed geometry of the matched rays (reference scene/gaussian_model.py:493 trains zval only)
        res = [None, None, None] + [d_ray[k] for k in _RAY[2:]] + [d_bg[k] for k in _BG]
        return tuple(r if need else None for r, need in zip(res, ctx.needs_input_grad))


def assemble(rayo=None, rayd=None, zval=None, scaling=None, rotation=None, opacity=None, features_dc=None,
             features_rest=None, bg_xyz=None, bg_scaling=None, bg_rotation=None, bg_opacity=None,
             bg_features_dc=None, bg_features_rest=None, with_sh=True):
    """(means3D [P,3], scales [P,3], rotations [P,4], opacities [P,1], shs [P,K,3]) of the hybrid model:
    the ray-based set (position rayo + rayd * zval) followed by the free set (position bg_xyz), activated as
    reference scene/gaussian_model.py:105-152 does.  Either set may be absent (all None)."""
    return _Assemble.apply(bool(with_sh), rayo, rayd, zval, scaling, rotation, opacity, features_dc, features_rest,
                           bg_xyz, bg_scaling, bg_rotation, bg_opacity, bg_features_dc, bg_features_rest)


def _attr(pc, name):
    t = getattr(pc, name, None)
    return t if isinstance(t, torch.Tensor) and t.dim() > 0 and t.shape[0] > 0 else None


def assemble_model(pc, with_sh=True):
    """`assemble` on a reference `GaussianModel` (attribute names of reference scene/gaussian_model.py:55-62 and
    the bg_* tensors tested at :109, :118, :126, :135, :146).  A model without ray attributes (plain 3DGS `_xyz`)
    is taken as one free set.  with_sh=False leaves the SH copy out (fifth output None): see `split_features`."""
    if _attr(pc, "_rayo") is None and _attr(pc, "_xyz") is not None and _attr(pc, "bg_xyz") is None:
        return assemble(bg_xyz=pc._xyz, bg_scaling=pc._scaling, bg_rotation=pc._rotation, bg_opacity=pc._opacity,
                        bg_features_dc=pc._features_dc, bg_features_rest=pc._features_rest, with_sh=with_sh)
    has_ray = _attr(pc, "_rayo") is not None
    has_bg = _attr(pc, "bg_xyz") is not None
    ray = dict(rayo=pc._rayo, rayd=pc._rayd, zval=pc._zval, scaling=pc._scaling, rotation=pc._rotation,
               opacity=pc._opacity, features_dc=pc._features_dc, features_rest=pc._features_rest) if has_ray else {}
    bg = dict(bg_xyz=pc.bg_xyz, bg_scaling=pc.bg_scaling, bg_rotation=pc.bg_rotation, bg_opacity=pc.bg_opacity,
              bg_features_dc=pc.bg_features_dc, bg_features_rest=pc.bg_features_rest) if has_bg else {}
    return assemble(**ray, **bg, with_sh=with_sh)


def split_features(pc):
    """(features_dc, features_rest, bg_features_dc, bg_features_rest) of the model when the operator can read them in
    place (16 SH coefficients, contiguous fp32: the `sh_split` input of scgaussian_b200.rasterizer), else None."""
    if _attr(pc, "_rayo") is None and _attr(pc, "_xyz") is not None and _attr(pc, "bg_xyz") is None:
        sets = (None, None, pc._features_dc, pc._features_rest)
    else:
        has_ray, has_bg = _attr(pc, "_rayo") is not None, _attr(pc, "bg_xyz") is not None
        sets = ((pc._features_dc, pc._features_rest) if has_ray else (None, None)) + \
               ((pc.bg_features_dc, pc.bg_features_rest) if has_bg else (None, None))
    for dc, rest in (sets[0:2], sets[2:4]):
        if dc is None:
            continue
        if rest is None or tuple(dc.shape[1:]) != (1, 3) or tuple(rest.shape[1:]) != (15, 3):
            return None
        if any(t.dtype != torch.float32 or not t.is_contiguous() or t.device.type != "cuda" for t in (dc, rest)):
            return None
    return sets


def add_densification_stats(pc, viewspace_point_tensor, update_filter=None, radii=None) -> None:
    """reference scene/gaussian_model.py:932-934 (`pc.add_densification_stats(viewspace_point_tensor, update_filter)`)
    and, when `radii` is given, reference train.py:192 (`max_radii2D[vis] = max(max_radii2D[vis], radii[vis])`) in
    the same launch -- in place on pc.xyz_gradient_accum / pc.denom / pc.max_radii2D, without the nonzero + host
    synchronisation each boolean-mask statement of the reference costs.  `update_filter=None` means `radii > 0`
    (what the reference passes)."""
    lib = _lib.load()
    grad = viewspace_point_tensor.grad
    if grad is None:
        raise ScgrError("add_densification_stats: viewspace_point_tensor has no .grad (call backward first)")
    if grad.device.type != "cuda":
        raise ScgrError("add_densification_stats runs on CUDA tensors only (no CPU path exists)")
    if update_filter is None and radii is None:
        raise ScgrError("add_densification_stats needs update_filter or radii")
    P = int(grad.shape[0])
    accum, denom = pc.xyz_gradient_accum, pc.denom
    max_radii = pc.max_radii2D if radii is not None else None
    state = [accum, denom] + ([max_radii] if max_radii is not None else [])
    for t in state:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.shape[0] != P or t.device != grad.device:
            raise ScgrError("add_densification_stats: statistics tensors must be contiguous fp32 [P, ...] on the gradient's device")
    grad = grad if grad.dtype == torch.float32 and grad.is_contiguous() else grad.float().contiguous()
    keep = [grad]
    if update_filter is not None:
        if update_filter.dtype != torch.bool or update_filter.shape[0] != P:
            raise ScgrError("add_densification_stats: update_filter must be a bool mask of length P")
        update_filter = update_filter.contiguous()
        keep.append(update_filter)
    if radii is not None:
        if radii.dtype != torch.int32 or radii.shape[0] != P:
            raise ScgrError("add_densification_stats: radii must be the operator's int32 [P] output")
        radii = radii.contiguous()
        keep.append(radii)
    with torch.cuda.device(grad.device):
        stream = C.c_void_p(torch.cuda.current_stream(grad.device).cuda_stream)
        check(lib.scgr_densification_stats(grad.data_ptr(), _p(update_filter), _p(radii), P, accum.data_ptr(),
                                           denom.data_ptr(), _p(max_radii), stream))


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """reference gaussian_renderer/__init__.py:20-118 with the model read through `assemble_model`: same
    arguments (minus the point-cloud dump switches :87-96, file I/O), same returned dict.  The two debug switches
    `pipe.compute_cov3D_python` / `pipe.convert_SHs_python` need the reference's own python helpers and are not on
    this path: use the reference's render() for them."""
    from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    if getattr(pipe, "compute_cov3D_python", False) or getattr(pipe, "convert_SHs_python", False):
        raise ScgrError("the fused render() has no python covariance / SH path: use the reference's render() for "
                        "pipe.compute_cov3D_python / pipe.convert_SHs_python")
    # SH degree 3 models (what the reference trains: arguments/__init__.py:49) hand their feature arrays to the operator
    # as they are; anything else goes through the assembled [P,K,3] copy
    split = split_features(pc) if override_color is None and os.environ.get("SCGR_SPLIT_SH", "1") != "0" else None
    means3D, scales, rotations, opacity, shs = assemble_model(pc, with_sh=split is None)
    # reference :28-32: the tensor whose .grad receives dL/dmean2D
    screenspace_points = torch.zeros_like(means3D, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=bool(getattr(pipe, "debug", False)))
    if split is not None:
        from .rasterizer import rasterize_gaussians_split
        rendered_image, radii, rendered_depth, rendered_alpha = rasterize_gaussians_split(
            means3D, screenspace_points, opacity, scales, rotations, split, raster_settings)
    else:
        rasterizer = GaussianRasterizer(raster_settings=raster_settings)
        rendered_image, radii, rendered_depth, rendered_alpha = rasterizer(
            means3D=means3D,
            means2D=screenspace_points,
            shs=shs if override_color is None else None,
            colors_precomp=override_color,
            opacities=opacity,
            scales=scales,
            rotations=rotations,
            cov3D_precomp=None)
    return {"render": rendered_image,
            "rendered_depth": rendered_depth,
            "rendered_alpha": rendered_alpha,
            "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0,
            "radii": radii}
