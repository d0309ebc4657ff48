"""Builds tests/emulation/_build/libemu_{model,preprocess}.so: scgaussian_b200/csrc/model.cu and preprocess.cu compiled
for the HOST through host_cuda_shim.h (TEST INFRASTRUCTURE -- see that header).  Source transformations: the launch
syntax,
    kernel<<<grid, threads, 0, L.stream>>>(args);   ->   emu_launch(grid, threads, [=] { kernel(args); });
the include of <cuda_runtime.h> in common.cuh (dropped: the shim stands in), and common.cuh's single inline-PTX
statement (`sqrt.approx.ftz.f32` -> sqrtf).
"""
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "scgaussian_b200", "csrc", "model.cu")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libemu_model.so")
_LAUNCH = re.compile(r"(\w+(?:<[^<>;]*>)?)<<<([^;]+?), (\w+), 0, L\.stream>>>\(\s*([^;]*)\);", re.S)
CSRC = os.path.join(ROOT, "scgaussian_b200", "csrc")
LIB_PRE = os.path.join(OUT_DIR, "libemu_preprocess.so")


def build() -> str:
    deps = [SRC, os.path.join(HERE, "host_cuda_shim.h"), os.path.join(HERE, "emu_model.cpp"), __file__,
            os.path.join(ROOT, "include", "scgr.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    body = open(SRC).read().replace('#include "common.cuh"', "")
    body, n = _LAUNCH.subn(lambda m: f"emu_launch(({m.group(2)}), ({m.group(3)}), [=] {{ {m.group(1)}({m.group(4)}); }});", body)
    assert n >= 9 and "<<<" not in body, f"launch rewrite incomplete ({n} launches rewritten)"
    with open(os.path.join(OUT_DIR, "model_body.inc"), "w") as f:
        f.write(body)
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-w", "-o", LIB,
                           os.path.join(HERE, "emu_model.cpp")], env=env, cwd=HERE)
    return LIB


def _rewrite_launches(text: str, at_least: int) -> str:
    text, n = _LAUNCH.subn(lambda m: f"emu_launch(({m.group(2)}), ({m.group(3)}), [=] {{ {m.group(1)}({m.group(4)}); }});", text)
    assert n >= at_least and "<<<" not in text, f"launch rewrite incomplete ({n} launches rewritten)"
    return text


def _compile(lib: str, cpp: str) -> None:
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-w", "-o", lib,
                           os.path.join(HERE, cpp)], env=env, cwd=HERE)


def build_preprocess() -> str:
    src, common = os.path.join(CSRC, "preprocess.cu"), os.path.join(CSRC, "common.cuh")
    deps = [src, common, os.path.join(HERE, "host_cuda_shim.h"), os.path.join(HERE, "emu_preprocess.cpp"), __file__,
            os.path.join(ROOT, "include", "scgr.h")]
    if os.path.exists(LIB_PRE) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PRE) for d in deps):
        return LIB_PRE
    os.makedirs(OUT_DIR, exist_ok=True)
    c = open(common).read().replace("#include <cuda_runtime.h>", "")
    c = c.replace('#include "../../include/scgr.h"', '#include "../../../include/scgr.h"')
    c, n = re.subn(r'asm\("sqrt\.approx\.ftz\.f32 %0, %1;"[^;]*;', "y = sqrtf(x);", c)
    assert n == 1 and "asm" not in c, "common.cuh: expected exactly one inline-PTX statement (sqrt.approx)"
    with open(os.path.join(OUT_DIR, "common_host.cuh"), "w") as f:
        f.write(c)
    body = open(src).read().replace('#include "common.cuh"', "")
    with open(os.path.join(OUT_DIR, "preprocess_body.inc"), "w") as f:
        f.write(_rewrite_launches(body, 8))
    _compile(LIB_PRE, "emu_preprocess.cpp")
    return LIB_PRE


def _common_host() -> None:
    c = open(os.path.join(CSRC, "common.cuh")).read().replace("#include <cuda_runtime.h>", "")
    c = c.replace('#include "../../include/scgr.h"', '#include "../../../include/scgr.h"')
    c, n = re.subn(r'asm\("sqrt\.approx\.ftz\.f32 %0, %1;"[^;]*;', "y = sqrtf(x);", c)
    assert n == 1 and "asm" not in c, "common.cuh: expected exactly one inline-PTX statement (sqrt.approx)"
    with open(os.path.join(OUT_DIR, "common_host.cuh"), "w") as f:
        f.write(c)


LIB_LOSS = os.path.join(OUT_DIR, "libemu_loss_knn.so")


def build_loss_knn() -> str:
    srcs = [os.path.join(CSRC, "loss.cu"), os.path.join(CSRC, "knn.cu")]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "host_cuda_shim.h"),
                   os.path.join(HERE, "emu_loss_knn.cpp"), __file__, os.path.join(ROOT, "include", "scgr.h")]
    if os.path.exists(LIB_LOSS) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_LOSS) for d in deps):
        return LIB_LOSS
    os.makedirs(OUT_DIR, exist_ok=True)
    _common_host()
    for src, name, n in ((srcs[0], "loss_body.inc", 3), (srcs[1], "knn_body.inc", 1)):
        body = open(src).read().replace('#include "common.cuh"', "")
        with open(os.path.join(OUT_DIR, name), "w") as f:
            f.write(_rewrite_launches(body, n))
    _compile(LIB_LOSS, "emu_loss_knn.cpp")
    return LIB_LOSS


if __name__ == "__main__":
    print(build(), build_preprocess(), build_loss_knn())
