"""`simple_knn._C.distCUDA2` drop-in (reference scene/gaussian_model.py:20, :444)."""
import numpy as np
import pytest
import torch

from oracle.knn_oracle import dist2_knn3


def test_oracle_closed_forms():
    # unit grid line: interior points have neighbours at 1, 1, 2 -> (1 + 1 + 4) / 3; the ends 1, 2, 3 -> (1 + 4 + 9) / 3
    line = np.stack([np.arange(6.0), np.zeros(6), np.zeros(6)], 1)
    got = dist2_knn3(line)
    assert np.allclose(got[[0, 5]], 14 / 3) and np.allclose(got[[2, 3]], 2.0) and np.allclose(got[[1, 4]], 2.0)
    # a duplicated point sees its twin at distance 0
    pts = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3]], float)
    assert np.allclose(dist2_knn3(pts)[0], (0 + 1 + 4) / 3)


def test_drop_in_module_has_no_cpu_path():
    from simple_knn._C import distCUDA2
    from scgaussian_b200 import ScgrError
    with pytest.raises(ScgrError, match="CUDA"):
        distCUDA2(torch.zeros(5, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 257, 3000])
def test_dist_cuda2_matches_oracle(n):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from simple_knn._C import distCUDA2
    g = torch.Generator().manual_seed(n)
    pts = torch.randn(n, 3, generator=g)
    if n >= 5:
        pts[3] = pts[1]                       # coincident points
    got = distCUDA2(pts.cuda()).cpu().numpy()
    want = dist2_knn3(pts.numpy())
    assert got.shape == (n,)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7), float(np.abs(got - want).max())
    # the reference's own use: clamp_min(dist2, 1e-7) then log(sqrt()) must be finite
    assert np.isfinite(np.log(np.sqrt(np.maximum(got, 1e-7)))).all()
