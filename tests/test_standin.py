"""The naive-CUDA stand-in baseline (baseline/standin: thread-per-Gaussian preprocess, cub 64-bit sort, 256-thread
tiles, 10 atomics per pair) against the CPU oracle: a third, independently structured implementation of SURVEY.md
Appendix A, and the thing bench.py times as `gpu_standin_baseline`.  It keeps the reference's un-culled (rect) instance
lists, so its R must equal the oracle's exactly."""
import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,W,H,deg,smed,yaw", [(3000, 131, 77, 3, 0.05, 15.0), (30000, 504, 378, 3, 0.03, 0.0),
                                                (1500, 200, 152, 2, 0.08, 0.0)])
def test_standin_matches_oracle(P, W, H, deg, smed, yaw):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from baseline.standin import standin as SB
    dev = torch.device("cuda:0")
    case = util.make_case(P, W, H, sh_degree=deg, scale_median=smed, w2c=O.yaw_w2c(yaw), bg=(0.1, 0.2, 0.3),
                          z_shift=-1.5 if P == 1500 else 0.0)
    grads = O.synth_upstream_grads(W, H)
    sb = SB.Standin(case, W, H, deg, dev)
    c, r, d, a = sb.forward(case, bg=case["bg"])
    g = sb.backward(*[x.to(dev).contiguous() for x in grads])
    torch.cuda.synchronize()
    co, (c2, r2, d2, a2), g2 = util.run_c_oracle(case, "f32", grads=grads)
    flips = util.flip_sets(co)
    n_rad = util.assert_radii_match("radii", r.cpu().numpy(), r2, flips)
    if n_rad == 0:
        assert sb.num_rendered == co.num_rendered          # same rect rule, no culling: identical instance count
    util.assert_image_close("color", c.cpu().numpy(), c2, flips)
    util.assert_image_close("depth", d.cpu().numpy(), d2, flips)
    util.assert_image_close("alpha", a.cpu().numpy(), a2, flips)
    for k in ("means3D", "means2D", "opacities", "shs", "scales", "rotations"):
        util.assert_grad_close(k, g[k].cpu().numpy(), g2[k].reshape(tuple(g[k].shape)), flips)
    SB.load().standin_release()
