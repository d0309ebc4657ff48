"""Host-side mirror of the photometric loss SCGaussian's training step computes (reference
train.py:160-161 with reference utils/loss_utils.py:40-41 `l1_loss` and :56-94 `ssim`) -- SURVEY.md
section 8(f) row f1.  Same names and argument meaning as the reference's functions; every byte of
compute goes through the C ABI of libscgr.so (include/scgr.h: scgr_photometric_forward /
scgr_photometric_backward).  No CPU / eager fallback.

    from scgaussian_b200.losses import l1_loss, ssim, photometric_loss
    loss = photometric_loss(image, gt_image, lambda_dssim)    # == (1-l)*l1_loss + l*(1-ssim), one kernel each way

Only what the reference's training step uses is on the fused path: `ssim(..., window_size=11,
size_average=True, mask=None)`; anything else raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ScgrError, check


def _planes(t: torch.Tensor):
    if t.dim() == 3:
        c, h, w = t.shape
    elif t.dim() == 4:
        c, h, w = t.shape[0] * t.shape[1], t.shape[2], t.shape[3]
    else:
        raise ScgrError("photometric loss expects [C,H,W] or [B,C,H,W] images")
    return int(c), int(h), int(w)


def _prep(t: torch.Tensor) -> torch.Tensor:
    if t.device.type != "cuda":
        raise ScgrError("the fused photometric loss runs on CUDA tensors only (no CPU path exists)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Photometric(torch.autograd.Function):
    """out3 = {Ll1, ssim, (1 - lambda) * Ll1 + lambda * (1 - ssim)}; `which` selects the returned entry."""

    @staticmethod
    def forward(ctx, image, gt, lambda_dssim: float, which: int):
        lib = _lib.load()
        # (fp32 / contiguity casts happen in _apply, outside the Function, so that autograd itself hands the
        # gradient back in the dtype of the caller's tensor)
        if ctx.needs_input_grad[1]:
            raise ScgrError("the fused photometric loss does not differentiate with respect to its target (the "
                            "reference's gt image is data, train.py:147); detach it or swap the arguments")
        if image.shape != gt.shape:
            raise ScgrError(f"image {tuple(image.shape)} and gt {tuple(gt.shape)} differ in shape")
        c, h, w = _planes(image)
        dev = image.device
        want_grad = bool(ctx.needs_input_grad[0])
        with torch.cuda.device(dev):
            scratch = torch.empty(lib.scgr_photometric_scratch_bytes(c, h, w), dtype=torch.uint8, device=dev)
            out3 = torch.empty(3, dtype=torch.float32, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(lib.scgr_photometric_forward(image.data_ptr(), gt.data_ptr(), c, h, w, float(lambda_dssim),
                                               scratch.data_ptr(), int(want_grad), out3.data_ptr(), stream))
        ctx.save_for_backward(image, gt, scratch)
        ctx.lambda_dssim = float(lambda_dssim)
        ctx.which = int(which)
        return out3[which]

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        image, gt, scratch = ctx.saved_tensors
        c, h, w = _planes(image)
        dev = image.device
        # d out3[which] / d image as a (lambda, upstream) pair for the one backward kernel:
        #   loss:  lambda = l, upstream = g;   Ll1: lambda = 0, upstream = g;   ssim: lambda = 1, upstream = -g
        lam = {0: 0.0, 1: 1.0, 2: ctx.lambda_dssim}[ctx.which]
        up = grad_out.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        if ctx.which == 1:
            up = -up
        with torch.cuda.device(dev):
            grad = torch.empty_like(image)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(lib.scgr_photometric_backward(image.data_ptr(), gt.data_ptr(), c, h, w, lam, scratch.data_ptr(),
                                                up.data_ptr(), grad.data_ptr(), stream))
        return grad, None, None, None


def _apply(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float, which: int) -> torch.Tensor:
    return _Photometric.apply(_prep(image), _prep(gt), lambda_dssim, which)


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2) -> torch.Tensor:
    """reference train.py:160-161: (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))."""
    return _apply(image, gt, lambda_dssim, 2)


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """reference utils/loss_utils.py:40-41."""
    return _apply(network_output, gt, 0.0, 0)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True, mask=None) -> torch.Tensor:
    """reference utils/loss_utils.py:56-94 for the arguments its training step uses."""
    if window_size != 11 or not size_average or mask is not None:
        raise ScgrError("fused ssim supports window_size=11, size_average=True, mask=None (what reference train.py:161 uses)")
    return _apply(img1, img2, 1.0, 1)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 5's loss path: the match-prior loss on the rendered depth, the DTU background term
# ---------------------------------------------------------------------------------------------------------------
class _MatchLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, table, keep, width: float, height: float):
        lib = _lib.load()
        H, W = int(depth.shape[-2]), int(depth.shape[-1])
        dev = depth.device
        with torch.cuda.device(dev):
            scratch = torch.empty(8, dtype=torch.float32, device=dev)
            out = torch.empty(1, dtype=torch.float32, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(lib.scgr_match_loss_forward(depth.data_ptr(), H, W, float(width), float(height), table.arr, table.n,
                                              scratch.data_ptr(), out.data_ptr(), stream))
        ctx.save_for_backward(depth, scratch)
        ctx.table, ctx.keep, ctx.size = table, keep, (float(width), float(height))
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        depth, scratch = ctx.saved_tensors
        H, W = int(depth.shape[-2]), int(depth.shape[-1])
        dev = depth.device
        up = grad_out.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            grad = torch.empty_like(depth)          # zero-filled by the library before the scatter
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(lib.scgr_match_loss_backward(depth.data_ptr(), H, W, ctx.size[0], ctx.size[1], ctx.table.arr, ctx.table.n,
                                               scratch.data_ptr(), up.data_ptr(), grad.data_ptr(), stream))
        return grad, None, None, None, None


def _match_table(view_gs: dict, img_name0: str, device):
    """The ScgrMatchPair table of one rendered view, from the reference model's own bookkeeping
    (`GaussianModel.view_gs`, reference scene/gaussian_model.py:284-366): built once per view and cached on the dict
    (the matches, rays and camera matrices are fixed for the whole training run)."""
    from ._lib import MATCH_MAX_PAIRS, ScgrMatchPair
    cache = view_gs.setdefault("__scgr_match_tables__", {})
    hit = cache.get(img_name0)
    if hit is not None:
        return hit
    pairs, keep = [], []
    f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()      # noqa: E731
    for img_name1, md in view_gs[img_name0]["match_infos"].items():
        other = view_gs[img_name1]
        mask0, mask1 = md["blender_mask"], other["match_infos"][img_name0]["blender_mask"]
        arrs = [f(md["uv"]), f(md["rays_o"]), f(md["rays_d"]), f(md["cam_rays_d"]),
                f(other["match_infos"][img_name0]["uv"]), f(mask0 * mask1)]
        keep += arrs
        w2c = other["w2c"].detach().float().cpu().reshape(4, 4)[:3].reshape(-1).tolist()
        intr = other["intr"].detach().float().cpu().reshape(-1).tolist()
        pairs.append(ScgrMatchPair(int(arrs[0].shape[0]), *[a.data_ptr() for a in arrs], (C.c_float * 12)(*w2c),
                                   (C.c_float * 9)(*intr)))
    if len(pairs) > MATCH_MAX_PAIRS:
        raise ScgrError(f"match loss: more than {MATCH_MAX_PAIRS} other views per view")
    table = (ScgrMatchPair * max(len(pairs), 1))(*pairs) if pairs else (ScgrMatchPair * 1)()
    table_len = len(pairs)
    hit = (table, table_len, keep)
    cache[img_name0] = hit
    return hit


class _Table:
    """ctypes array + its true length (an empty view has a 1-element placeholder array)."""

    def __init__(self, arr, n):
        self.arr, self.n = arr, n


def get_matchloss_from_renderdepth(pc, cam0, depth0: torch.Tensor, loss_state=None) -> torch.Tensor:
    """reference scene/gaussian_model.py:241-282 `GaussianModel.get_matchloss_from_renderdepth(cam0, depth0, loss_state)`
    (called at reference train.py:164 on `render_pkg["rendered_depth"]`): same arguments with the model first, same
    value, differentiable in depth0 -- one kernel forward, one backward, for all the other views of cam0 at once.
    `loss_state` is unused, as in the reference."""
    if depth0.device.type != "cuda":
        raise ScgrError("the fused match loss runs on CUDA tensors only (no CPU path exists)")
    name0 = cam0.image_name
    vg = pc.view_gs[name0]
    table, n, keep = _match_table(pc.view_gs, name0, depth0.device)
    d = depth0.squeeze(0) if depth0.dim() == 3 else depth0
    if d.dim() != 2:
        raise ScgrError("match loss expects the [1,H,W] rendered depth")
    d = d if d.dtype == torch.float32 else d.float()
    return _MatchLoss.apply(d.contiguous(), _Table(table, n), keep, float(vg["width"]), float(vg["height"]))


def dtu_background_mask(gt_image: torch.Tensor, threshold: float = 30.0 / 255.0, window: int = 50) -> torch.Tensor:
    """reference train.py:150-158: `bg_mask` = pixels whose brightest channel is below `threshold` and whose `window - 1`
    upper neighbours in the column are too (the 49-step shift loop), returned as bool [1,H,W]; `gt_image` [C,H,W] is
    zeroed IN PLACE where masked, as the reference does.  One launch instead of ~100."""
    lib = _lib.load()
    if gt_image.device.type != "cuda":
        raise ScgrError("dtu_background_mask runs on CUDA tensors only (no CPU path exists)")
    if gt_image.dim() != 3 or gt_image.dtype != torch.float32 or not gt_image.is_contiguous():
        raise ScgrError("dtu_background_mask expects a contiguous fp32 [C,H,W] image (it is modified in place)")
    c, h, w = (int(x) for x in gt_image.shape)
    dev = gt_image.device
    with torch.cuda.device(dev):
        mask = torch.empty(1, h, w, dtype=torch.bool, device=dev)
        count = torch.empty(1, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib.scgr_bg_mask(gt_image.data_ptr(), c, h, w, float(threshold), int(window), mask.data_ptr(),
                               count.data_ptr(), stream))
    return mask


class _MaskedMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, mask):
        lib = _lib.load()
        dev = values.device
        n = values.numel()
        with torch.cuda.device(dev):
            scratch = torch.empty(lib.scgr_masked_mean_scratch_bytes(n), dtype=torch.uint8, device=dev)
            out2 = torch.empty(2, dtype=torch.float32, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(lib.scgr_masked_mean_forward(values.data_ptr(), mask.data_ptr(), n, scratch.data_ptr(), out2.data_ptr(),
                                               stream))
        ctx.save_for_backward(mask, out2)
        ctx.shape = values.shape
        return out2[0]

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        mask, out2 = ctx.saved_tensors
        dev = mask.device
        up = grad_out.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            grad = torch.empty(ctx.shape, dtype=torch.float32, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(lib.scgr_masked_mean_backward(mask.data_ptr(), mask.numel(), out2.data_ptr(), up.data_ptr(),
                                                grad.data_ptr(), stream))
        return grad, None


def masked_mean(values: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """`values[mask].mean()` (reference train.py:168 on `render_pkg["rendered_alpha"][bg_mask]`) without the boolean-mask
    gather (a nonzero + host synchronisation in torch); differentiable in `values`."""
    if values.device.type != "cuda":
        raise ScgrError("masked_mean runs on CUDA tensors only (no CPU path exists)")
    if mask.dtype != torch.bool or mask.numel() != values.numel():
        raise ScgrError("masked_mean expects a bool mask with as many elements as `values`")
    return _MaskedMean.apply(_prep(values), mask.contiguous())
