"""CPU oracle of BASELINE config 5's loss path (TEST INFRASTRUCTURE: only tests/ may import this; the product never
does).  Restates, dtype-generic (run it in float64 for the truth):

    reference scene/gaussian_model.py:241-282   GaussianModel.get_matchloss_from_renderdepth
    reference train.py:151-158                  bg_mask (threshold + the 49-step shift loop), gt zeroing
    reference train.py:167-168                  rendered_alpha[bg_mask].mean()

PINNED: tests/golden/prior_golden.npz holds outputs (values and autograd gradients) of the reference's own method /
statements run in the build container (tests/golden/make_prior_golden.py);
tests/test_prior.py::test_oracle_matches_reference_golden checks this file against them.
"""
import numpy as np
import torch


def bilinear_zeros(img, fx, fy):
    """F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=False) at un-normalised coordinates (fx, fy)."""
    H, W = img.shape
    x0, y0 = torch.floor(fx), torch.floor(fy)
    out = torch.zeros_like(fx)
    for dx in (0, 1):
        for dy in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            w = (1 - (fx - xi).abs()) * (1 - (fy - yi).abs())
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            v = img[yi.clamp(0, H - 1).long(), xi.clamp(0, W - 1).long()]
            out = out + torch.where(ok, v * w, torch.zeros_like(v))
    return out


def match_loss(depth, pairs, width, height):
    """depth [H,W]; pairs: list of dicts with uv0 [n,2], rays_o, rays_d, cam_rays_d [n,3], uv1 [n,2], valid [n]
    (mask0 * mask1), w2c1 [4,4], intr1 [3,3] -- reference :245-279, one loop iteration per pair."""
    H, W = depth.shape
    total = depth.new_zeros(())
    for p in pairs:
        uv0 = p["uv0"]
        nx, ny = uv0[:, 0] / width * 2 - 1, uv0[:, 1] / height * 2 - 1
        d = bilinear_zeros(depth, ((nx + 1) * W - 1) / 2, ((ny + 1) * H - 1) / 2)
        zval = (d / p["cam_rays_d"][:, 2]).unsqueeze(-1)
        world = (p["rays_o"] + p["rays_d"] * zval).T                                  # [3, n]
        cam = (p["w2c1"] @ torch.cat([world, torch.ones_like(world[:1])]))[:3]
        xyz = p["intr1"] @ cam
        xy = xyz[:2] / (xyz[2:] + 1e-8)
        inside = ((xy[0] > 0) & (xy[0] < width) & (xy[1] > 0) & (xy[1] < height)).to(depth.dtype)
        valid = (p["valid"] > 0).to(depth.dtype)
        scale = torch.tensor([width, height], dtype=depth.dtype).reshape(2, 1)
        cur = ((xy - p["uv1"].T).abs() / scale).mean(dim=0)
        total = total + (cur * inside * valid).sum() / ((inside * valid).sum() + 1e-8)
    return total


def pairs_from_golden(G, name0, dtype=torch.float64):
    names = [str(n) for n in G["match_names"]]
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)      # noqa: E731
    pairs = []
    for name1 in names:
        if name1 == name0:
            continue
        k = f"view_{name0}_{name1}_"
        pairs.append(dict(uv0=t(G[k + "uv"]), rays_o=t(G[k + "rays_o"]), rays_d=t(G[k + "rays_d"]),
                          cam_rays_d=t(G[k + "cam_rays_d"]), uv1=t(G[f"view_{name1}_{name0}_uv"]),
                          valid=t(G[k + "blender_mask"]) * t(G[f"view_{name1}_{name0}_blender_mask"]),
                          w2c1=t(G[f"view_{name1}_w2c"]), intr1=t(G[f"view_{name1}_intr"])))
    return pairs


def dtu_background_mask(gt, threshold=30.0 / 255.0, window=50):
    """gt [C,H,W] numpy -> (mask bool [1,H,W], gt zeroed where masked).  A pixel is masked when it and the `window - 1`
    pixels above it in its column (those that exist) are all darker than `threshold`."""
    gt = np.array(gt, copy=True)
    bg = gt.max(0) < threshold                                # [H, W]
    H = bg.shape[0]
    run = np.zeros_like(bg, dtype=np.int64)
    for y in range(H):
        run[y] = np.where(bg[y], (run[y - 1] if y else 0) + 1, 0)
    need = np.minimum(np.arange(H) + 1, window)[:, None]
    mask = run >= need
    gt[:, mask] = 0.0
    return mask[None], gt


def masked_mean(values, mask):
    return values[mask].mean()
