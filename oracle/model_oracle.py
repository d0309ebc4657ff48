"""CPU oracle of the two elementwise passes around the rasterizer in a training step (SURVEY.md section 8f rows f2,
f3).  TEST INFRASTRUCTURE: only tests/ may import this; the product path (scgaussian_b200/model.py, optim.py)
computes in libscgr.so and has no CPU route.

PINNED: both functions are checked against tests/golden/model_golden.npz, produced by tests/golden/
make_model_golden.py running the reference's own `GaussianModel` properties / `training_setup` optimizers in this
container (tests/test_model.py::test_oracle_matches_reference_golden).

`assemble` restates reference scene/gaussian_model.py:105-152 with the same torch primitives the reference calls
(so autograd yields the backward); `adam_step` restates torch/optim/adam.py `_single_tensor_adam` -- the optimizer
the reference constructs at scene/gaussian_model.py:502, :512 -- in numpy.
"""
import numpy as np
import torch
import torch.nn.functional as F


def assemble(ray: dict, bg: dict, dtype=torch.float64):
    """ray: rayo, rayd, zval, scaling, rotation, opacity, features_dc, features_rest (or {}); bg: xyz, scaling,
    rotation, opacity, features_dc, features_rest (or {}).  Returns (means3D, scales, rotations, opacities, shs) and
    the leaf tensors (dict name -> tensor with requires_grad) so that a test can call autograd on them."""
    leaves = {}

    def leaf(prefix, d, k, grad=True):
        t = torch.as_tensor(np.asarray(d[k])).to(dtype).clone().requires_grad_(grad)
        leaves[prefix + k] = t
        return t

    xyz, scal, rot, opa, dc, rest = [], [], [], [], [], []
    if ray:
        rayo, rayd = leaf("ray_", ray, "rayo", False), leaf("ray_", ray, "rayd", False)
        xyz.append(rayo + rayd * leaf("ray_", ray, "zval"))                      # reference :124
        prefix, d = "ray_", ray
        scal.append(torch.exp(leaf(prefix, d, "scaling")))                        # :108
        rot.append(F.normalize(leaf(prefix, d, "rotation")))                      # :118
        opa.append(torch.sigmoid(leaf(prefix, d, "opacity")))                     # :145
        dc.append(leaf(prefix, d, "features_dc"))
        rest.append(leaf(prefix, d, "features_rest"))
    if bg:
        xyz.append(leaf("bg_", bg, "xyz"))                                        # :126-127
        prefix, d = "bg_", bg
        scal.append(torch.exp(leaf(prefix, d, "scaling")))                        # :110-111
        rot.append(F.normalize(leaf(prefix, d, "rotation")))                      # :119-120
        opa.append(torch.sigmoid(leaf(prefix, d, "opacity")))                     # :147-148
        dc.append(leaf(prefix, d, "features_dc"))
        rest.append(leaf(prefix, d, "features_rest"))
    shs = torch.cat((torch.cat(dc), torch.cat(rest)), dim=1)                      # :136-140
    return (torch.cat(xyz), torch.cat(scal), torch.cat(rot), torch.cat(opa), shs), leaves


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float32):
    """One `torch.optim.Adam` update (no weight decay, no amsgrad), `step` counted from 1.  Returns (p, m, v)."""
    p, g, m, v = (np.asarray(a, dtype=dtype) for a in (p, g, m, v))
    f = dtype
    m = m + (g - m) * f(1.0 - beta1)                                  # exp_avg.lerp_(grad, 1 - beta1)
    v = v * f(beta2) + f(1.0 - beta2) * g * g                         # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / f(bc2 ** 0.5) + f(eps)
    p = p - f(step_size) * (m / denom)                                # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p.astype(dtype), m.astype(dtype), v.astype(dtype)


def densification_stats(grad2d, update_filter, radii, accum, denom, max_radii):
    """reference train.py:192 + scene/gaussian_model.py:932-934 on numpy arrays; returns the updated copies."""
    f = np.asarray(update_filter, dtype=bool)
    accum, denom = np.array(accum, dtype=np.float32), np.array(denom, dtype=np.float32)
    max_radii = np.array(max_radii, dtype=np.float32)
    if radii is not None:
        max_radii[f] = np.maximum(max_radii[f], np.asarray(radii)[f].astype(np.float32))        # train.py:192
    g = np.asarray(grad2d, dtype=np.float32)
    accum[f] += np.sqrt(g[f, 0:1] * g[f, 0:1] + g[f, 1:2] * g[f, 1:2])                          # :933
    denom[f] += 1                                                                               # :934
    return accum, denom, max_radii
