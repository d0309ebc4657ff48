"""CPU oracle of `simple_knn._C.distCUDA2` (TEST INFRASTRUCTURE: only tests/ may import this).

The reference calls it once (scene/gaussian_model.py:444: `dist2 = clamp_min(distCUDA2(points), 1e-7)`); the
implementation lives in the external, un-pinned `simple-knn` package (reference README.md:24, SURVEY.md section 2b
row N8) which is absent here, so this restates its published definition -- for every point the mean of the squared
distances to its 3 nearest OTHER points -- by brute force in float64.  PARITY UNPINNED by any reference-owned
vector; anchored on closed-form cases (tests/test_knn.py)."""
import numpy as np


def dist2_knn3(points: np.ndarray) -> np.ndarray:
    p = np.asarray(points, dtype=np.float64)
    n = p.shape[0]
    out = np.zeros(n)
    for i in range(n):
        d = ((p - p[i]) ** 2).sum(1)
        d = np.delete(d, i)
        k = np.sort(d)[:3]
        out[i] = k.mean() if len(k) else 0.0
    return out
