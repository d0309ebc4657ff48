"""Cross-process determinism probe (GPU box): one config-3 forward, prints R and a hash of every
intermediate (records, tiles_touched, depth order, instance list, ranges, image) on one line.
Run it N times and diff the lines."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from scgaussian_b200 import synthetic as O  # noqa: E402
from scgaussian_b200 import GaussianRasterizationSettings  # noqa: E402
from scgaussian_b200 import rasterizer as R  # noqa: E402

P, W, H = 1_000_000, 1920, 1080
dev = torch.device("cuda:0")
cam = O.make_camera(W, H)
sc = O.synth_scene(P, W, H, sh_degree=3, scale_median=0.01, seed=0)
t = {k: v.to(dev).contiguous() for k, v in sc.items()}
s = GaussianRasterizationSettings(H, W, cam["tanfovx"], cam["tanfovy"], torch.zeros(3, device=dev), 1.0,
                                  cam["viewmatrix"].to(dev), cam["projmatrix"].to(dev), 3, cam["campos"].to(dev), False, False)
args = (t["means3D"], t["opacities"], t["shs"], None, t["scales"], t["rotations"], None)


def h(x):
    return hashlib.sha1(x.contiguous().cpu().numpy().tobytes()).hexdigest()[:8]


for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    color, radii, depth, alpha, state = R.rasterize_forward_raw(*args, s)
    torch.cuda.synchronize()
    dv = R.debug_views(state, P, s)
    print(f"R={state.num_rendered} in={h(t['means3D'])}{h(t['scales'])}{h(t['rotations'])} rec={h(dv['record'][:, :8])} rgb={h(dv['record'][:, 8:11])} "
          f"radii={h(radii)} tt={h(dv['tiles_touched'])} order={h(dv['depth_order'])} pl={h(dv['point_list'])} "
          f"ranges={h(dv['ranges'])} color={h(color)}", flush=True)
