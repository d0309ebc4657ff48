"""Work counters of the render kernels (GPU box).  Builds a SEPARATE library with -DSCGR_STATS
(scgaussian_b200/libscgr_stats.so; the product library compiles none of the counters), runs one
forward + backward of BASELINE config 3 through it and prints what the warps actually did:
list entries scanned, entries that reached the blend loop, 8x4 slots evaluated, contributing
pixel pairs, reductions / atomics in the backward, plus the n_contrib distribution.

    python tools/render_stats.py [P W H scale_median]      (build only: --build)
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from scgaussian_b200 import build as B  # noqa: E402

STATS_LIB = os.path.join(os.path.dirname(B.LIB), "libscgr_stats.so")


def build_stats_lib():
    srcs = [os.path.join(B.CSRC, s) for s in B.SOURCES]
    if os.path.exists(STATS_LIB) and all(os.path.getmtime(s) < os.path.getmtime(STATS_LIB) for s in srcs + B.HEADERS):
        return STATS_LIB
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, "-DSCGR_STATS", "-o", STATS_LIB, *srcs], env=env)
    return STATS_LIB


def main():
    build_stats_lib()
    if "--build" in sys.argv:
        print(STATS_LIB)
        return
    import torch
    from scgaussian_b200 import _lib
    _lib.LIB_PATH = STATS_LIB
    from scgaussian_b200 import GaussianRasterizationSettings
    from scgaussian_b200 import rasterizer as R
    from scgaussian_b200 import synthetic as O

    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    P, W, H = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (1_000_000, 1920, 1080)
    smed = float(a[3]) if len(a) >= 4 else 0.01
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    cam = O.make_camera(W, H)
    sc = O.synth_scene(P, W, H, sh_degree=3, scale_median=smed, seed=0)
    t = {k: v.to(dev).contiguous() for k, v in sc.items()}
    gC, gD, gA = [g.to(dev).contiguous() for g in O.synth_upstream_grads(W, H, seed=1)]
    s = GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], torch.zeros(3, device=dev), 1.0,
                                      cam["viewmatrix"].to(dev), cam["projmatrix"].to(dev), 3, cam["campos"].to(dev),
                                      False, False)
    args_in = (t["means3D"], t["opacities"], t["shs"], None, t["scales"], t["rotations"], None)
    raw = C.CDLL(STATS_LIB)
    raw.scgr_debug_render_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    buf = (C.c_ulonglong * 32)()
    color, radii, depth, alpha, state = R.rasterize_forward_raw(*args_in, s)
    R.rasterize_backward_raw(state, *args_in, s, gC, gD, gA)
    torch.cuda.synchronize()
    assert raw.scgr_debug_render_stats(buf, 1) == 0
    v = list(buf)
    Rn = state.num_rendered
    dv = R.debug_views(state, P, s)
    nc = dv["n_contrib"].float()
    tiles_y, tiles_x = (H + 15) // 16, (W + 15) // 16
    pad = torch.zeros(tiles_y * 16, tiles_x * 16, device=nc.device)
    pad[:H, :W] = nc
    tmax = pad.view(tiles_y, 16, tiles_x, 16).amax(dim=(1, 3))
    rng = dv["ranges"].long()
    tlen = (rng[:, 1] - rng[:, 0]).clamp(min=0).float()
    out = {
        "P": P, "W": W, "H": H, "R": Rn,
        "fwd": {"work_items": v[6], "list_entries_total": v[7], "batches": v[0], "scanned": v[1], "hit": v[2],
                "slots": v[3], "cand_pairs": v[4], "go_pairs": v[5],
                "slots_per_hit": v[3] / max(v[2], 1), "pairs_per_slot": v[5] / max(v[3], 1) ,
                "hit_frac_of_scanned": v[2] / max(v[1], 1), "scanned_frac_of_R": v[1] / max(Rn, 1)},
        "bwd": {"work_items": v[15], "toDo_total": v[16], "batches": v[8], "scanned": v[9], "hit": v[10],
                "slots": v[11], "ok_pairs": v[12], "reductions": v[13], "atomics": v[14],
                "slots_per_hit": v[11] / max(v[10], 1), "pairs_per_slot": v[12] / max(v[11], 1),
                "hit_frac_of_scanned": v[10] / max(v[9], 1)},
        "n_contrib": {"mean": float(nc.mean()), "p50": float(nc.median()), "max": float(nc.max()),
                      "tile_max_mean": float(tmax.mean()), "tile_len_mean": float(tlen.mean()),
                      "tile_len_max": float(tlen.max())},
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
