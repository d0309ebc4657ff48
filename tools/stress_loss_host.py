"""Randomised shapes (strip / band / vector-width edge cases, band-count overrides) through the streaming loss kernels run ON THE
HOST (tests/emulation) against the float64 oracle; meant to be run under AddressSanitizer:
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 SCGR_EMU_ASAN=1 python tools/stress_loss_host.py [seed] [cases]
Development tool: test infrastructure only (imports oracle/ through tests/)."""
import os, sys, ctypes as C, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.emulation import build
from tests import test_kernel_emulation as T
from tests import test_loss as TL
lib = C.CDLL(build.build_loss_knn()); lib.emu_photometric_scratch_bytes.restype = C.c_size_t
os.environ["SCGR_LOSS_VARIANT"] = "1"
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
special_w = [1, 2, 3, 4, 5, 6, 10, 11, 12, 115, 116, 117, 120, 121, 122, 126, 127, 128, 231, 232, 233, 236, 348]
special_h = [1, 2, 5, 6, 10, 11, 12, 32, 33, 34, 43, 44, 45, 65, 66, 67, 98, 99, 100, 110, 131]
worst = 0.0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    c = random.choice([1, 1, 2, 3])
    w = random.choice(special_w) if random.random() < 0.6 else random.randint(1, 300)
    h = random.choice(special_h) if random.random() < 0.6 else random.randint(1, 160)
    bands = random.choice([None, None, 1, 2, 3, 5, 50])
    if bands is None: os.environ.pop("SCGR_LOSS_BANDS", None)
    else: os.environ["SCGR_LOSS_BANDS"] = str(bands)
    gen = torch.Generator().manual_seed(it * 7919 + w * 31 + h)
    gt = torch.rand(c, h, w, generator=gen)
    img = (gt + 0.2 * torch.randn(c, h, w, generator=gen)).clamp(0, 1)
    o = TL._oracle_all(img.numpy(), gt.numpy(), torch.float64)
    ll1, s, loss, g = T._host_loss(lib, img.numpy(), gt.numpy(), 0.2)
    ev = max(abs(a - b) for a, b in zip((ll1, s, loss), o[:3]))
    eg = TL._rel(g, o[3])
    assert np.isfinite(g).all(), (c, h, w, bands)
    assert ev < 5e-6 and eg < TL.RTOL_GRAD, (c, h, w, bands, ev, eg)
    worst = max(worst, eg)
    print(f"{it:3d} C={c} H={h} W={w} bands={bands}: values {ev:.1e} grad rel {eg:.1e}", flush=True)
print("all ok; worst gradient error", worst)
