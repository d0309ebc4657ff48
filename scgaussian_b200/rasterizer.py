"""Host-side mirror of the operator surface SCGaussian imports at
reference gaussian_renderer/__init__.py:15 (`from diff_gaussian_rasterization import
GaussianRasterizationSettings, GaussianRasterizer`) and uses at :38-53 and :100-108.

Same names, same argument meaning, same return order ``(color, radii, depth, alpha)``, same error
behaviour as the external package the reference's README
installs -- but every byte of compute goes through the C ABI of libscgr.so (include/scgr.h) via
ctypes: torch only owns the memory and the stream.  There is no CPU / eager fallback; a missing
library raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _lib
from ._lib import ScgrGaussians, ScgrGrads, ScgrView, ScgrDebugViews, ScgrError, check


class GaussianRasterizationSettings(NamedTuple):
    """The 12 fields the reference passes by keyword (gaussian_renderer/__init__.py:38-51)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# --------------------------------------------------------------------------------------------
# scratch / capacity management
# --------------------------------------------------------------------------------------------
# How the forward learns R (the instance count) to size the binning buffer:
#   "fused"      (default) one scgr_forward() call: the binning buffer is pre-sized from the previous
#                view's R (+25 %); the library enqueues both stages (stage 2 reads R on the device and
#                refuses to run past the buffer) and waits for R on a zero-copy pinned word only to
#                report which case it was; a view that outgrew the headroom is finished with an
#                exactly sized buffer (scgr_forward_render), stage 1 being kept
#   "sync"       the reference's protocol in two calls: blocking read of R, exactly sized buffer
#   "optimistic" both stages enqueued blind with the pre-sized buffer, one validation sync at the end
#   "async"      (opt-in) NO host wait at all: both stages are enqueued blind with a generously pre-sized buffer
#                (ASYNC_HEADROOM x the largest recent R) and the call returns at once, so the host can run a whole
#                iteration ahead of the GPU -- what the small-scene regime needs (504x378: ~260 us of GPU work per
#                step against ~300 us of host work; with the reference's blocking read of R the two add up instead of
#                overlapping).  R is picked up later from the pinned status words and only steers the next buffer
#                size; ForwardState.num_rendered is -1.  The price: a view whose R outgrows the headroom cannot be
#                re-rendered -- its images are NaN, its gradients zero (the kernels refuse to run past the buffer),
#                `dropped_views` counts it and a warning is printed.  With the default headroom of 2x that takes a
#                view with twice the instances of every recent one.
_BINNING_MODE = os.environ.get("SCGR_BINNING", "fused")
ASYNC_HEADROOM = float(os.environ.get("SCGR_ASYNC_HEADROOM", "2.0"))
dropped_views = 0          # async mode: views that outgrew their binning buffer (images NaN, gradients zero)
_capacity_hint = {}        # device index -> last num_rendered
_pinned_status = {}        # (device index, thread id) -> pinned int64[2]
launch_counter = 0         # number of libscgr stage calls (bench.py reports kernels from this)
need_capacity_count = 0    # forwards whose pre-sized binning buffer was too small (finished by a second stage-2 call)


def _status_buffer(device: torch.device) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), threading.get_ident())
    buf = _pinned_status.get(key)
    if buf is None:
        buf = torch.zeros(4, dtype=torch.int64).pin_memory()     # [0:2] {R, overflow} of stage 1 / fused, [2:4] of stage 2 (async mode)
        _pinned_status[key] = buf
    return buf


def _scratch(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator hands out >= 512-byte aligned blocks
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _make_view(s: GaussianRasterizationSettings, device, keep: list) -> ScgrView:
    bg = _f32c(s.bg, device)
    vm = _f32c(s.viewmatrix, device)
    pm = _f32c(s.projmatrix, device)
    cp = _f32c(s.campos, device)
    keep += [bg, vm, pm, cp]
    return ScgrView(int(s.image_height), int(s.image_width), float(s.tanfovx), float(s.tanfovy),
                    bg.data_ptr(), float(s.scale_modifier), vm.data_ptr(), pm.data_ptr(),
                    int(s.sh_degree), cp.data_ptr(), int(bool(s.prefiltered)), int(bool(s.debug)))


def _make_gaussians(means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, sh_split=None) -> ScgrGaussians:
    P = int(means3D.shape[0])
    M = int(sh.shape[1]) if sh is not None and sh.numel() > 0 else 0
    g = ScgrGaussians(P, M, _ptr(means3D), _ptr(opacities), _ptr(sh), _ptr(colors_precomp),
                      _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp))
    if sh_split is not None:
        dc0, rest0, dc1, rest1 = sh_split
        g.sh_coeffs = 16
        g.sh_dc[0], g.sh_rest[0] = _ptr(dc0), _ptr(rest0)
        g.sh_dc[1], g.sh_rest[1] = _ptr(dc1), _ptr(rest1)
        g.sh_n0 = 0 if dc0 is None else int(dc0.shape[0])
    return g


def _check_split(sh_split, P, device):
    """sh_split = (features_dc, features_rest) of the two sets of the hybrid model (reference scene/gaussian_model.py:
    131-140), either set possibly None: contiguous fp32 [n,1,3] / [n,15,3] on `device`, n0 + n1 == P."""
    if len(sh_split) != 4:
        raise ScgrError("sh_split = (features_dc, features_rest, bg_features_dc, bg_features_rest)")
    n = 0
    for dc, rest in (sh_split[0:2], sh_split[2:4]):
        if (dc is None) != (rest is None):
            raise ScgrError("sh_split: features_dc and features_rest of a set go together")
        if dc is None:
            continue
        if tuple(dc.shape[1:]) != (1, 3) or tuple(rest.shape[1:]) != (15, 3) or rest.shape[0] != dc.shape[0]:
            raise ScgrError("sh_split needs [n,1,3] / [n,15,3] arrays (16 SH coefficients)")
        for t in (dc, rest):
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                raise ScgrError("sh_split arrays must be contiguous fp32 tensors on the operator's device")
        n += int(dc.shape[0])
    if n != P:
        raise ScgrError(f"sh_split holds {n} Gaussians, means3D {P}")


class ForwardState(NamedTuple):
    """What the matching backward needs (the reference saves geomBuffer / binningBuffer /
    imgBuffer / num_rendered the same way)."""
    geometry: torch.Tensor
    binning: torch.Tensor
    image: torch.Tensor
    capacity: int
    num_rendered: int
    radii: Optional[torch.Tensor] = None


def rasterize_forward_raw(means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp,
                          s: GaussianRasterizationSettings, sh_split=None):
    """Runs the two forward stages of libscgr on the current stream.  Inputs must be contiguous fp32
    CUDA tensors (or None).  `sh_split` (instead of `sh`): the SH coefficients as the hybrid model stores them, see
    _check_split.  Returns (color, radii, depth, alpha, ForwardState)."""
    global launch_counter, need_capacity_count
    lib = _lib.load()
    device = means3D.device
    if device.type != "cuda":
        raise ScgrError("the rasterizer runs on CUDA tensors only (no CPU path exists)")
    P = int(means3D.shape[0])
    H, W = int(s.image_height), int(s.image_width)
    keep: list = []
    with torch.cuda.device(device):
        view = _make_view(s, device, keep)
        if sh_split is not None:
            _check_split(sh_split, P, device)
        g = _make_gaussians(means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, sh_split)
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        geometry = _scratch(lib.scgr_geometry_bytes(P), device)
        image = _scratch(lib.scgr_image_bytes(W, H), device)
        radii = torch.empty(P, dtype=torch.int32, device=device)
        color = torch.empty(3, H, W, dtype=torch.float32, device=device)
        depth = torch.empty(1, H, W, dtype=torch.float32, device=device)
        alpha = torch.empty(1, H, W, dtype=torch.float32, device=device)
        status = _status_buffer(device)
        key = device.index

        def run_render(capacity: int, want_status: bool = False) -> torch.Tensor:
            # `status` is the shared pinned word scgr_forward() spins on.  An asynchronous {R, overflow} copy into
            # it that is still in flight when the NEXT forward stores its sentinel would land on top of the
            # sentinel and be taken for that view's R (no host sync separates two forwards): it is therefore
            # requested only by the one protocol that synchronises the stream right after ("optimistic").
            global launch_counter
            binning = _scratch(lib.scgr_binning_bytes(P, W, H, capacity), device)
            check(lib.scgr_forward_render(C.byref(view), C.byref(g), geometry.data_ptr(), binning.data_ptr(),
                                          capacity, image.data_ptr(), color.data_ptr(), depth.data_ptr(),
                                          alpha.data_ptr(), status.data_ptr() if want_status else None, stream))
            launch_counter += 1
            return binning

        hint = _capacity_hint.get(key) if _BINNING_MODE in ("optimistic", "fused", "async") else None
        if hint is not None and _BINNING_MODE == "async":
            global dropped_views
            # what the GPU has reported so far (an EARLIER view's counts: nothing here waits for this one's)
            seen, overflowed = int(status[2]), int(status[3])
            if overflowed:
                status[3] = 0
                dropped_views += 1
                print(f"scgaussian_b200: async binning dropped a view ({seen} instances did not fit; SCGR_ASYNC_HEADROOM)",
                      flush=True)
            hint = max(seen, int(status[0]), int(hint * 0.98))
            capacity = int(hint * ASYNC_HEADROOM) + 65536
            check(lib.scgr_forward_geometry(C.byref(view), C.byref(g), geometry.data_ptr(), radii.data_ptr(),
                                            status.data_ptr(), stream))
            binning = _scratch(lib.scgr_binning_bytes(P, W, H, capacity), device)
            check(lib.scgr_forward_render(C.byref(view), C.byref(g), geometry.data_ptr(), binning.data_ptr(), capacity,
                                          image.data_ptr(), color.data_ptr(), depth.data_ptr(), alpha.data_ptr(),
                                          status.data_ptr() + 16, stream))
            launch_counter += 2
            _capacity_hint[key] = hint
            return color, radii, depth, alpha, ForwardState(geometry, binning, image, capacity, -1, radii)
        if hint is not None and _BINNING_MODE == "fused":
            capacity = max(int(hint * 1.25) + 4096, 4096)
            binning = _scratch(lib.scgr_binning_bytes(P, W, H, capacity), device)
            rc = lib.scgr_forward(C.byref(view), C.byref(g), geometry.data_ptr(), radii.data_ptr(),
                                  binning.data_ptr(), capacity, image.data_ptr(), color.data_ptr(),
                                  depth.data_ptr(), alpha.data_ptr(), status.data_ptr(), stream)
            launch_counter += 1
            if rc == _lib.NEED_CAPACITY:      # this view outgrew the headroom: stage 1 is done, R is known
                need_capacity_count += 1
                R = int(status[0])
                capacity = R
                binning = run_render(capacity)
            else:
                check(rc)
                R = int(status[0])
        elif hint is not None:
            # both stages back to back, one validation sync at the end (GPU never idles mid-forward)
            capacity = max(int(hint * 1.25) + 4096, 4096)
            check(lib.scgr_forward_geometry(C.byref(view), C.byref(g), geometry.data_ptr(), radii.data_ptr(),
                                            None, stream))
            launch_counter += 1
            binning = run_render(capacity, want_status=True)
            torch.cuda.current_stream(device).synchronize()
            R, overflow = int(status[0]), int(status[1])
            if overflow or R > capacity:
                need_capacity_count += 1
                capacity = R
                binning = run_render(capacity)
        else:
            # the reference's protocol: one blocking read of num_rendered between the stages
            check(lib.scgr_forward_geometry(C.byref(view), C.byref(g), geometry.data_ptr(), radii.data_ptr(),
                                            status.data_ptr(), stream))
            launch_counter += 1
            torch.cuda.current_stream(device).synchronize()
            R = int(status[0])
            capacity = R
            binning = run_render(capacity)
        _capacity_hint[key] = R
    return color, radii, depth, alpha, ForwardState(geometry, binning, image, capacity, R, radii)


def rasterize_backward_raw(state: ForwardState, means3D, opacities, sh, colors_precomp, scales, rotations,
                           cov3Ds_precomp, s: GaussianRasterizationSettings, grad_color, grad_depth, grad_alpha,
                           out: Optional[dict] = None, accumulate: bool = False, sh_split=None) -> dict:
    """Runs scgr_backward.  `out` may hold pre-allocated gradient tensors (e.g. views into one flat
    all-reduce buffer, scgaussian_b200/parallel.py); missing ones are allocated.  Every returned
    tensor is fully written by the kernels.  out["stats"] ([P,2], optional) receives the two per-view
    densification terms of reference scene/gaussian_model.py:932-934, out["live"] ([P], optional) 1 / 0 per Gaussian that
    did / did not receive any gradient (what the row-sparse all-reduce keys on).  accumulate=True ADDS the parameter
    gradients (and the statistics) to what `out` holds -- gradient accumulation over the views a rank renders
    before the batch's single all-reduce; it needs every gradient tensor to be passed in."""
    global launch_counter
    lib = _lib.load()
    device = means3D.device
    P = int(means3D.shape[0])
    H, W = int(s.image_height), int(s.image_width)
    out = dict(out) if out else {}

    missing = []

    def need(name, ref, shape):
        if ref is None or ref.numel() == 0:
            return None
        t = out.get(name)
        if t is None:
            missing.append(name)
            t = torch.empty(shape, dtype=torch.float32, device=device)
            out[name] = t
        assert t.is_contiguous() and t.dtype == torch.float32 and tuple(t.shape) == tuple(shape), name
        return t

    keep: list = []
    with torch.cuda.device(device):
        gm3 = need("means3D", means3D, (P, 3))
        if "means2D" not in out:
            out["means2D"] = torch.empty(P, 3, dtype=torch.float32, device=device)
        gm2 = out["means2D"]
        gsh = need("shs", sh, tuple(sh.shape) if sh is not None else ())
        gcol = need("colors_precomp", colors_precomp, (P, 3))
        gop = need("opacities", opacities, tuple(opacities.shape))
        gsc = need("scales", scales, (P, 3))
        grot = need("rotations", rotations, (P, 4))
        gcov = need("cov3D_precomp", cov3Ds_precomp, (P, 6))
        gsplit = [None] * 4
        if sh_split is not None:      # gradients in the model's own layout: "sh_split0".."sh_split3"
            gsplit = [need(f"sh_split{k}", t, tuple(t.shape) if t is not None else ()) for k, t in enumerate(sh_split)]
        stats = out.get("stats")
        if stats is not None:
            assert stats.is_contiguous() and stats.dtype == torch.float32 and tuple(stats.shape) == (P, 2), "stats"
            if state.radii is None:
                raise ScgrError("densification statistics need the forward's radii (ForwardState.radii)")
        live = out.get("live")
        if live is not None:
            assert live.is_contiguous() and live.dtype == torch.float32 and live.numel() == P, "live"
        if accumulate and missing:
            raise ScgrError(f"accumulate=True needs the gradient tensors to add into: {missing} not in `out`")
        if P == 0:
            return out
        view = _make_view(s, device, keep)
        g = _make_gaussians(means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, sh_split)
        grads = ScgrGrads(_ptr(gm3), _ptr(gm2), _ptr(gsh), _ptr(gcol), _ptr(gop), _ptr(gsc), _ptr(grot), _ptr(gcov),
                          _ptr(stats), _ptr(state.radii) if stats is not None else None, int(bool(accumulate)),
                          _ptr(live))
        grads.dL_dsh_dc[0], grads.dL_dsh_rest[0] = _ptr(gsplit[0]), _ptr(gsplit[1])
        grads.dL_dsh_dc[1], grads.dL_dsh_rest[1] = _ptr(gsplit[2]), _ptr(gsplit[3])
        gc = _f32c(grad_color, device)
        gd = _f32c(grad_depth, device)
        ga = _f32c(grad_alpha, device)
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        check(lib.scgr_backward(C.byref(view), C.byref(g), state.geometry.data_ptr(), state.binning.data_ptr(),
                                state.capacity, state.image.data_ptr(), gc.data_ptr(), gd.data_ptr(),
                                ga.data_ptr(), C.byref(grads), stream))
        launch_counter += 1
    return out


def _snapshot(path: str, payload) -> None:
    """Reference `debug` behaviour: dump the inputs that made the extension fail."""
    try:
        torch.save(tuple(p.detach().cpu().clone() if isinstance(p, torch.Tensor) else p for p in payload), path)
        print(f"\nAn error occured in the rasterizer. Writing {path} for debugging.")
    except Exception:
        pass


# Where the autograd nodes below put their gradients.  None: fresh tensors.  A callable returning a dict of pre-allocated
# tensors (scgaussian_b200.parallel.FlatGradBuffer.capture()): scgr_backward writes the parameter gradients, the two
# densification statistics and the live counts straight into them -- for a data-parallel trainer, straight into the flat
# all-reduce buffer, with no packing copies between loss.backward() and the collective.
class _Sink:      # process-wide, not thread-local: autograd runs a CUDA node's backward on its own engine thread
    fn = None


_grad_sink = _Sink()


def set_gradient_sink(fn) -> None:
    _grad_sink.fn = fn


def _sink_out() -> Optional[dict]:
    fn = getattr(_grad_sink, "fn", None)
    return fn() if fn is not None else None


class _RasterizeGaussians(torch.autograd.Function):
    """same argument order as the external package's autograd Function."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        args = (means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp, raster_settings)
        try:
            color, radii, depth, alpha, state = rasterize_forward_raw(*args)
        except Exception:
            if raster_settings.debug:
                _snapshot("snapshot_fw.dump", args[:-1])
            raise
        ctx.raster_settings = raster_settings
        ctx.state = state
        ctx.num_rendered = state.num_rendered
        ctx.save_for_backward(means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        s = ctx.raster_settings
        means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp = ctx.saved_tensors
        H, W = int(s.image_height), int(s.image_width)
        dev = means3D.device
        if grad_color is None:
            grad_color = torch.zeros(3, H, W, device=dev)
        if grad_depth is None:
            grad_depth = torch.zeros(1, H, W, device=dev)
        if grad_alpha is None:
            grad_alpha = torch.zeros(1, H, W, device=dev)
        try:
            g = rasterize_backward_raw(ctx.state, means3D, opacities, sh, colors_precomp, scales, rotations,
                                       cov3Ds_precomp, s, grad_color, grad_depth, grad_alpha, out=_sink_out())
        except Exception:
            if s.debug:
                _snapshot("snapshot_bw.dump", (means3D, opacities, sh, colors_precomp, scales, rotations,
                                               cov3Ds_precomp, grad_color, grad_depth, grad_alpha))
            raise
        return (g.get("means3D"), g.get("means2D"), g.get("shs"), g.get("colors_precomp"), g.get("opacities"),
                g.get("scales"), g.get("rotations"), g.get("cov3D_precomp"), None)


class _RasterizeGaussiansSplit(torch.autograd.Function):
    """The operator with the SH coefficients taken straight from the hybrid model's four arrays (SURVEY.md section 8f row
    f2, second half; include/scgr.h: ScgrGaussians.sh_dc / sh_rest): same outputs, the gradients of the coefficients come
    back in the arrays' own layout.  Scale / rotation path only (what the training step uses)."""

    @staticmethod
    def forward(ctx, means3D, means2D, opacities, scales, rotations, dc0, rest0, dc1, rest1, raster_settings):
        split = (dc0, rest0, dc1, rest1)
        color, radii, depth, alpha, state = rasterize_forward_raw(means3D, opacities, None, None, scales, rotations, None,
                                                                  raster_settings, sh_split=split)
        ctx.raster_settings = raster_settings
        ctx.state = state
        ctx.num_rendered = state.num_rendered
        ctx.has = [t is not None for t in split]
        ctx.save_for_backward(means3D, opacities, scales, rotations, *[t for t in split if t is not None])
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        s = ctx.raster_settings
        saved = list(ctx.saved_tensors)
        means3D, opacities, scales, rotations = saved[:4]
        it = iter(saved[4:])
        split = tuple(next(it) if h else None for h in ctx.has)
        H, W = int(s.image_height), int(s.image_width)
        dev = means3D.device
        if grad_color is None:
            grad_color = torch.zeros(3, H, W, device=dev)
        if grad_depth is None:
            grad_depth = torch.zeros(1, H, W, device=dev)
        if grad_alpha is None:
            grad_alpha = torch.zeros(1, H, W, device=dev)
        g = rasterize_backward_raw(ctx.state, means3D, opacities, None, None, scales, rotations, None, s, grad_color,
                                   grad_depth, grad_alpha, sh_split=split)
        return (g.get("means3D"), g.get("means2D"), g.get("opacities"), g.get("scales"), g.get("rotations"),
                g.get("sh_split0"), g.get("sh_split1"), g.get("sh_split2"), g.get("sh_split3"), None)


def rasterize_gaussians_split(means3D, means2D, opacities, scales, rotations, sh_split, raster_settings):
    return _RasterizeGaussiansSplit.apply(means3D, means2D, opacities, scales, rotations, *sh_split, raster_settings)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def _prep(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    """None / empty -> None; otherwise contiguous fp32 on `device` (differentiably)."""
    if t is None or t.numel() == 0:
        return None
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class GaussianRasterizer(nn.Module):
    """SURVEY.md section 8a row a5.  Usage is exactly the reference's
    (gaussian_renderer/__init__.py:53,100-108)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            lib = _lib.load()
            s = self.raster_settings
            device = positions.device
            if device.type != "cuda":
                raise ScgrError("markVisible runs on CUDA tensors only")
            pos = _prep(positions, device)
            P = int(positions.shape[0])
            present = torch.zeros(P, dtype=torch.uint8, device=device)
            if P > 0:
                vm = _f32c(s.viewmatrix, device)
                with torch.cuda.device(device):
                    stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
                    check(lib.scgr_mark_visible(pos.data_ptr(), P, vm.data_ptr(), present.data_ptr(), stream))
            return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        s = self.raster_settings
        none = lambda t: t is None or (isinstance(t, torch.Tensor) and t.numel() == 0 and t.dim() <= 1)
        if (none(shs) and none(colors_precomp)) or (not none(shs) and not none(colors_precomp)):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((none(scales) or none(rotations)) and none(cov3D_precomp)) or \
                ((not none(scales) or not none(rotations)) and not none(cov3D_precomp)):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise ScgrError("means3D must have dimensions (num_points, 3)")
        device = means3D.device
        return rasterize_gaussians(
            _prep(means3D, device) if means3D.numel() else means3D.float(), means2D, _prep(shs, device),
            _prep(colors_precomp, device), _prep(opacities, device) if opacities.numel() else opacities.float(),
            _prep(scales, device), _prep(rotations, device), _prep(cov3D_precomp, device), s)


def mark_visible(positions, viewmatrix, projmatrix=None):
    """Functional form of GaussianRasterizer.markVisible (`_C.mark_visible` of the external package)."""
    s = GaussianRasterizationSettings(0, 0, 1.0, 1.0, torch.zeros(3), 1.0, viewmatrix,
                                      viewmatrix if projmatrix is None else projmatrix, 0, torch.zeros(3),
                                      False, False)
    return GaussianRasterizer(s).markVisible(positions)


def debug_views(state: ForwardState, P: int, s: GaussianRasterizationSettings) -> dict:
    """Copies the intermediate state of a forward out of the scratch buffers (tests only)."""
    lib = _lib.load()
    H, W = int(s.image_height), int(s.image_width)
    dv = ScgrDebugViews()
    check(lib.scgr_debug_views(P, W, H, state.capacity, state.geometry.data_ptr(), state.binning.data_ptr(),
                               state.image.data_ptr(), C.byref(dv)))

    def grab(buf: torch.Tensor, addr, nbytes, dtype):
        off = int(addr) - buf.data_ptr()
        assert 0 <= off and off + nbytes <= buf.numel(), (off, nbytes, buf.numel())
        return buf[off:off + nbytes].clone().view(dtype)

    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    R = state.num_rendered
    out = {
        "record": grab(state.geometry, dv.record, P * 48, torch.float32).view(P, 12),
        "tiles_touched": grab(state.geometry, dv.tiles_touched, P * 4, torch.int32),
        "depth_order": grab(state.geometry, dv.depth_order, P * 4, torch.int32),
        "status": grab(state.geometry, dv.num_rendered, 16, torch.int64),
        "point_list": grab(state.binning, dv.point_list, R * 4, torch.int32),
        "ranges": grab(state.binning, dv.ranges, tiles * 8, torch.int32).view(tiles, 2),
        "n_contrib": grab(state.image, dv.n_contrib, W * H * 4, torch.int32).view(H, W),
        "final_T": grab(state.image, dv.final_T, W * H * 4, torch.float32).view(H, W),
    }
    out["ranges"][out["ranges"][:, 1] == 0] = 0      # empty tiles are stored as (0xffffffff, 0)
    return out
