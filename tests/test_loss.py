"""Photometric loss: reference train.py:160-161 =
(1 - l) * l1_loss + l * (1 - ssim), reference utils/loss_utils.py:40-41, 56-94.

CPU: the oracle (oracle/loss_oracle.py) is PINNED to the reference's own functions through
tests/golden/loss_golden.npz (made by tests/golden/make_loss_golden.py, which imports
/root/reference/utils/loss_utils.py).  GPU: the fused CUDA kernels (scgaussian_b200/csrc/loss.cu,
called through the C ABI) against the oracle and against the golden vectors directly."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as LO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.npz")
TOL_VAL = 2e-6      # fp32 values: |a - b| (all of l1, ssim, loss are O(1))
RTOL_GRAD = 2e-4    # gradients: max|a - b| / max|b|


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _oracle_all(img, gt, dtype):
    x = torch.tensor(img, dtype=dtype, requires_grad=True)
    y = torch.tensor(gt, dtype=dtype)
    ll1, s = LO.l1_loss(x, y), LO.ssim(x, y)
    loss = LO.photometric_loss(x, y, 0.2)
    g_loss, = torch.autograd.grad(loss, x, retain_graph=True)
    g_ssim, = torch.autograd.grad(s, x, retain_graph=True)
    g_l1, = torch.autograd.grad(ll1, x)
    return float(ll1.detach()), float(s.detach()), float(loss.detach()), g_loss.numpy(), g_ssim.numpy(), g_l1.numpy()


@pytest.mark.parametrize("name", ["a", "b", "c"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_reference_golden(name, dtype):
    g = np.load(GOLD)
    ll1, s, loss, g_loss, g_ssim, g_l1 = _oracle_all(g[f"{name}_img"], g[f"{name}_gt"], dtype)
    assert abs(ll1 - float(g[f"{name}_l1"])) < TOL_VAL
    assert abs(s - float(g[f"{name}_ssim"])) < TOL_VAL
    assert abs(loss - float(g[f"{name}_loss"])) < TOL_VAL
    assert _rel(g_loss, g[f"{name}_g_loss"]) < RTOL_GRAD
    assert _rel(g_ssim, g[f"{name}_g_ssim"]) < RTOL_GRAD
    assert _rel(g_l1, g[f"{name}_g_l1"]) < 1e-6


def test_window_is_the_references():
    w = LO.window_1d()
    assert w.dtype == torch.float32 and w.numel() == 11
    assert abs(float(w.sum()) - 1.0) < 1e-6 and float(w[5]) == float(w.max()) and torch.equal(w, w.flip(0))


def test_host_logic_rejects_what_is_not_on_the_fused_path():
    from scgaussian_b200 import losses
    from scgaussian_b200._lib import ScgrError
    x = torch.rand(3, 8, 8)
    with pytest.raises(ScgrError):
        losses.photometric_loss(x, x)                       # CPU tensors: no CPU path exists
    with pytest.raises(ScgrError):
        losses.ssim(x, x, window_size=7)
    with pytest.raises(ScgrError):
        losses.ssim(x, x, size_average=False)
    with pytest.raises(ScgrError):
        losses.ssim(x, x, mask=torch.ones(8, 8))


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


@pytest.fixture(params=["tiles", "stream"])
def variant(request, monkeypatch):
    """The two kernel designs of scgaussian_b200/csrc/loss.cu (SCGR_LOSS_VARIANT, read on every launch): 32x32 tiles and
    streaming column strips.  Every GPU test of the fused loss runs under both."""
    monkeypatch.setenv("SCGR_LOSS_VARIANT", {"tiles": "0", "stream": "1"}[request.param])
    return request.param


def _fused_all(img, gt, dev, lam=0.2):
    from scgaussian_b200 import losses
    x = torch.tensor(img, dtype=torch.float32, device=dev, requires_grad=True)
    y = torch.tensor(gt, dtype=torch.float32, device=dev)
    ll1, s = losses.l1_loss(x, y), losses.ssim(x, y)
    loss = losses.photometric_loss(x, y, lam)
    g_loss, = torch.autograd.grad(loss, x)
    g_ssim, = torch.autograd.grad(s, x)
    g_l1, = torch.autograd.grad(ll1, x)
    return (float(ll1.detach()), float(s.detach()), float(loss.detach()), g_loss.cpu().numpy(), g_ssim.cpu().numpy(),
            g_l1.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_fused_loss_matches_reference_golden(dev, name, variant):
    g = np.load(GOLD)
    ll1, s, loss, g_loss, g_ssim, g_l1 = _fused_all(g[f"{name}_img"], g[f"{name}_gt"], dev)
    assert abs(ll1 - float(g[f"{name}_l1"])) < TOL_VAL
    assert abs(s - float(g[f"{name}_ssim"])) < TOL_VAL
    assert abs(loss - float(g[f"{name}_loss"])) < TOL_VAL
    assert _rel(g_loss, g[f"{name}_g_loss"]) < RTOL_GRAD
    assert _rel(g_ssim, g[f"{name}_g_ssim"]) < RTOL_GRAD
    assert _rel(g_l1, g[f"{name}_g_l1"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 1, 1), (3, 5, 7), (1, 16, 16), (3, 17, 33), (2, 3, 40, 24), (3, 378, 504), (1, 70, 131),
                                   (2, 45, 116), (1, 100, 250)])
def test_fused_loss_matches_oracle_f64(dev, shape, variant):
    g = torch.Generator().manual_seed(sum(shape))
    gt = torch.rand(*shape, generator=g)
    img = (gt + 0.2 * torch.randn(*shape, generator=g)).clamp(0, 1)
    c = shape[-3] * (shape[0] if len(shape) == 4 else 1)
    fold = lambda t: t.reshape(c, shape[-2], shape[-1]).numpy()
    o = _oracle_all(fold(img), fold(gt), torch.float64)
    f = _fused_all(img.numpy(), gt.numpy(), dev, lam=0.2)
    for k in range(3):
        assert abs(f[k] - o[k]) < 5e-6, (k, f[k], o[k])
    for k in range(3, 6):
        assert _rel(f[k].reshape(o[k].shape), o[k]) < RTOL_GRAD, k


@pytest.mark.gpu
def test_fused_loss_full_size_properties(dev, variant):
    """1080p (BASELINE config 3 image size): identities that need no oracle."""
    from scgaussian_b200 import losses
    g = torch.Generator().manual_seed(3)
    y = torch.rand(3, 1080, 1920, generator=g).to(dev)
    x = (y + 0.1 * torch.randn(3, 1080, 1920, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
    # identical images: Ll1 = 0, ssim = 1, loss = 0
    assert float(losses.l1_loss(y, y)) == 0.0
    assert abs(float(losses.ssim(y, y)) - 1.0) < 1e-6
    assert abs(float(losses.photometric_loss(y, y, 0.2))) < 1e-6
    # loss is the stated combination of its two parts; deterministic; l1 agrees with torch
    ll1, s, loss = losses.l1_loss(x, y), losses.ssim(x, y), losses.photometric_loss(x, y, 0.2)
    assert abs(float(loss) - (0.8 * float(ll1) + 0.2 * (1.0 - float(s)))) < 1e-6
    assert float(losses.photometric_loss(x, y, 0.2)) == float(loss)
    assert abs(float(ll1) - float((x - y).abs().mean())) < 1e-6
    # the upstream scalar scales the gradient linearly; gradient of the parts adds up
    g1, = torch.autograd.grad(loss, x, retain_graph=True)
    g3, = torch.autograd.grad(3.0 * losses.photometric_loss(x, y, 0.2), x)
    assert float((g3 - 3.0 * g1).abs().max()) <= 1e-6 * float(g1.abs().max()) * 3
    gl, = torch.autograd.grad(ll1, x)
    gs, = torch.autograd.grad(s, x)
    assert float((g1 - (0.8 * gl - 0.2 * gs)).abs().max()) <= 2e-6 * float(g1.abs().max())
    # no_grad forward works and matches
    with torch.no_grad():
        assert float(losses.photometric_loss(x, y, 0.2)) == float(loss)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 1080, 1920), (3, 2160, 3840), (3, 1081, 1918)])
def test_fused_loss_variants_agree_at_full_size(dev, shape, monkeypatch):
    """The two kernel designs sum the same 121 products per moment in different orders: values to rounding of the mean,
    gradients to 1e-5 of their scale, at the image sizes of BASELINE configs 3 and 4 and at one with ragged strips
    (W % 4 != 0: the element-wise row path)."""
    from scgaussian_b200 import losses
    g = torch.Generator().manual_seed(shape[1])
    y = torch.rand(*shape, generator=g).to(dev)
    x = (y + 0.1 * torch.randn(*shape, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
    out = {}
    for v in ("0", "1"):
        monkeypatch.setenv("SCGR_LOSS_VARIANT", v)
        loss = losses.photometric_loss(x, y, 0.2)
        gr, = torch.autograd.grad(loss, x)
        out[v] = (float(loss), gr)
    assert abs(out["0"][0] - out["1"][0]) < 2e-6
    assert float((out["0"][1] - out["1"][1]).abs().max()) <= 1e-5 * float(out["0"][1].abs().max())
    assert bool(torch.isfinite(out["1"][1]).all())
