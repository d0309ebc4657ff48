"""CPU oracle of the photometric loss (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this; the product never does).

Plain-torch restatement, dtype-generic (run it in float64 for the truth), of
    reference utils/loss_utils.py:40-41   l1_loss
    reference utils/loss_utils.py:46-54   gaussian / create_window (11 taps, sigma 1.5, float32 window)
    reference utils/loss_utils.py:56-94   ssim / _ssim (mask=None, size_average=True)
    reference train.py:160-161            loss = (1 - l) * Ll1 + l * (1 - ssim)
PINNED: tests/golden/loss_golden.npz holds outputs (values and autograd gradients) of the reference's
own functions imported from /root/reference (tests/golden/make_loss_golden.py);
tests/test_loss.py::test_oracle_matches_reference_golden checks this file against them.
"""
import math

import torch
import torch.nn.functional as F


def window_1d(window_size=11, sigma=1.5):
    # reference :46-48 -- built in float32 exactly as the reference does (torch.Tensor of python floats)
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


def l1_loss(x, y):
    return (x - y).abs().mean()


def ssim(img1, img2, window_size=11):
    c = img1.shape[-3]
    w1 = window_1d(window_size).unsqueeze(1)
    win = (w1 @ w1.t()).float()[None, None].expand(c, 1, window_size, window_size).contiguous().to(img1.dtype)
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, win, padding=pad, groups=c)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s11 = conv(img1 * img1) - mu1_sq
    s22 = conv(img2 * img2) - mu2_sq
    s12 = conv(img1 * img2) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s11 + s22 + C2))
    return m.mean()


def photometric_loss(image, gt, lambda_dssim=0.2):
    return (1.0 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1.0 - ssim(image, gt))
