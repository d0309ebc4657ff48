"""Builds tests/emulation/_build/libemu_model.so: scgaussian_b200/csrc/model.cu compiled for the HOST through
host_cuda_shim.h (TEST INFRASTRUCTURE -- see that header).  The only source transformation is the launch syntax:
    kernel<<<grid, threads, 0, L.stream>>>(args);   ->   emu_launch(grid, threads, [=] { kernel(args); });
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
_LAUNCH = re.compile(r"(\w+(?:<\d+>)?)<<<([^,]+), ([^,]+), 0, L\.stream>>>\(\s*([^;]*)\);")


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


if __name__ == "__main__":
    print(build())
