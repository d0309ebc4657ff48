"""Generates tests/golden/loss_golden.npz by IMPORTING the reference's own loss code
(/root/reference/utils/loss_utils.py: l1_loss :40-41, ssim :56-94) and running it, with autograd, on
small seeded images.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_loss_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from utils.loss_utils import l1_loss, ssim  # noqa: E402


def main():
    out = {}
    g = torch.Generator().manual_seed(7)
    for name, (c, h, w) in {"a": (3, 37, 53), "b": (3, 16, 16), "c": (1, 9, 70)}.items():
        gt = torch.rand(c, h, w, generator=g)
        img = (gt + 0.15 * torch.randn(c, h, w, generator=g)).clamp(0, 1)
        img[:, : h // 3] = gt[:, : h // 3]                 # a region with |x - y| == 0 exactly
        img.requires_grad_(True)
        ll1 = l1_loss(img, gt)
        s = ssim(img, gt)
        loss = (1.0 - 0.2) * ll1 + 0.2 * (1.0 - s)         # reference train.py:161, lambda_dssim = 0.2
        g_loss, = torch.autograd.grad(loss, img, retain_graph=True)
        g_ssim, = torch.autograd.grad(s, img, retain_graph=True)
        g_l1, = torch.autograd.grad(ll1, img)
        out[f"{name}_img"], out[f"{name}_gt"] = img.detach().numpy(), gt.numpy()
        out[f"{name}_l1"], out[f"{name}_ssim"], out[f"{name}_loss"] = ll1.item(), s.item(), loss.item()
        out[f"{name}_g_loss"], out[f"{name}_g_ssim"], out[f"{name}_g_l1"] = g_loss.numpy(), g_ssim.numpy(), g_l1.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("a_")})


if __name__ == "__main__":
    main()
