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
