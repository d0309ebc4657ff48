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

# SCGR_EMU_ASAN=1: AddressSanitizer build (a memcheck of the kernels on the host).  Run the tests as
#   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 SCGR_EMU_ASAN=1 \
#       python -m pytest tests/test_kernel_emulation.py -p no:cacheprovider
# so that torch's own allocations carry red zones as well: an access past the end of any tensor aborts with the
# offending line of the kernel source.
ASAN = os.environ.get("SCGR_EMU_ASAN") == "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "scgaussian_b200", "csrc", "model.cu")
OUT_DIR = os.path.join(HERE, "_build")
_SUFFIX = "_asan.so" if ASAN else ".so"
LIB = os.path.join(OUT_DIR, "libemu_model" + _SUFFIX)
_LAUNCH = re.compile(r"(\w+(?:<[^<>;]*>)?)<<<([^;]+?), (\w+), 0, L\.stream>>>\(\s*([^;]*)\);", re.S)
CSRC = os.path.join(ROOT, "scgaussian_b200", "csrc")
LIB_PRE = os.path.join(OUT_DIR, "libemu_preprocess" + _SUFFIX)


def build() -> str:
    deps = [SRC, os.path.join(HERE, "host_cuda_shim.h"), os.path.join(HERE, "emu_model.cpp"), __file__,
            os.path.join(ROOT, "include", "scgr.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    body = open(SRC).read().replace('#include "common.cuh"', "")
    body = _replace_fn_body(body, "sqrt_approx", "return sqrtf(x);")       # model.cu's two inline-PTX helpers (Adam)
    body = _replace_fn_body(body, "div_approx", "return a / b;")
    assert "asm" not in body
    with open(os.path.join(OUT_DIR, "model_body.inc"), "w") as f:
        f.write(_rewrite_launches(body, 9))
    _compile(LIB, "emu_model.cpp")
    return LIB


def _rewrite_launches(text: str, at_least: int) -> str:
    """NAME<TEMPLATE ARGS><<<grid, threads, smem, L.stream>>>(ARGS)  ->  emu_launch((grid), (threads), [=] { NAME<..>(ARGS); })
    with balanced parentheses, so that launches inside macros and over several lines are handled as well."""
    out, pos, n = [], 0, 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            break
        # kernel name (+ template arguments) right before <<<
        j = k
        if text[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                depth += text[j] == ">"
                depth -= text[j] == "<"
                if depth == 0:
                    break
        i = j
        while text[i - 1].isalnum() or text[i - 1] == "_":
            i -= 1
        name = text[i:k]
        e = text.index(">>>", k)
        cfg = [c.strip() for c in text[k + 3:e].rsplit(",", 3)]          # grid, threads, smem, stream
        assert len(cfg) == 4 and cfg[3] == "L.stream", cfg
        a = text.index("(", e)
        depth, b = 0, a
        while True:
            depth += text[b] == "("
            depth -= text[b] == ")"
            if depth == 0:
                break
            b += 1
        out.append(text[pos:i])
        out.append(f"emu_launch(({cfg[0]}), ({cfg[1]}), [=] {{ {name}({text[a + 1:b]}); }})")
        pos = b + 1
        n += 1
    out.append(text[pos:])
    text = "".join(out)
    assert n >= at_least and "<<<" not in text, f"launch rewrite incomplete ({n} launches rewritten)"
    return text


def _compile(lib: str, cpp: str) -> None:
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    extra = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if ASAN else []
    subprocess.check_call([gxx, "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-w", *extra, "-o", lib,
                           os.path.join(HERE, cpp)], env=env, cwd=HERE)


_VOLATILE = [   # binning.cu's four inline-PTX statements: volatile global loads / stores of the look-back words
    (r'asm volatile\("ld\.volatile\.global\.u64 %0, \[%1\];"[^;]*;', "v = *(volatile const unsigned long long*)p;"),
    (r'asm volatile\("st\.volatile\.global\.u64 \[%0\], %1;"[^;]*;', "*(volatile unsigned long long*)p = v;"),
    (r'asm volatile\("ld\.volatile\.global\.u32 %0, \[%1\];"[^;]*;', "v = *(volatile const uint32_t*)p;"),
    (r'asm volatile\("st\.volatile\.global\.u32 \[%0\], %1;"[^;]*;', "*(volatile uint32_t*)p = v;"),
]


def _replace_fn_body(text: str, name: str, body: str) -> str:
    """Replaces the body of the (unique) function definition `... name(args) { ... }`."""
    m = re.search(r"\b" + re.escape(name) + r"\([^)]*\)\s*\{", text)
    assert m, name
    a = m.end() - 1
    depth, b = 0, a
    while True:
        depth += text[b] == "{"
        depth -= text[b] == "}"
        if depth == 0:
            break
        b += 1
    return text[:a] + "{ " + body + " }" + text[b + 1:]


# render.cu's inline PTX, function by function.  The TMA plumbing becomes a synchronous copy: the data is in the
# staging buffer when bulk_g2s returns, so the mbarrier protocol has nothing left to wait for.
_RENDER_PTX = {
    "ex2": "return exp2f(x);",
    "fma2": "return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));",      # fma.rn.f32x2: two IEEE fused ops
    "mul2": "return make_float2(a.x * b.x, a.y * b.y);",
    "add2": "return make_float2(a.x + b.x, a.y + b.y);",
    "rcp_approx": "return 1.0f / x;",
    "gate_pair": "return (pos < lc && power <= 0.f && og_raw >= 1.0f / 255.0f) ? og_raw : 0.f;",
}
# common.cuh's TMA plumbing (render.cu and preprocess.cu both stage through it): a synchronous copy behind a real
# mbarrier phase protocol (host_cuda_shim.h: emu_mbar_*)
_TMA_PTX = {
    "smem_u32": "(void)p; return 0u;",
    "mbar_init": "emu_mbar_init(bar, count);",
    "mbar_fence_init": "",
    "fence_proxy_async": "",
    "cp_async16": "std::memcpy(dst, src, 16);",
    "cp_async_commit": "",
    "cp_async_wait_all_but_one": "",
    "cp_async_wait_all": "",
    "mbar_arrive": "emu_mbar_update(bar, -1, 0);",
    "mbar_expect_tx": "emu_mbar_update(bar, -1, (long long)bytes);",
    "bulk_g2s": "std::memcpy(dst, src, bytes); emu_mbar_update(bar, 0, -(long long)bytes);",
    "mbar_wait": "emu_mbar_wait(bar, parity);",
    "pdl_wait": "",                # programmatic dependent launch: kernels run one after the other here
    "pdl_trigger": "",
}


def build_preprocess() -> str:
    """preprocess.cu + binning.cu + render.cu: the whole operator, forward and backward."""
    srcs = [os.path.join(CSRC, "preprocess.cu"), os.path.join(CSRC, "binning.cu"), os.path.join(CSRC, "render.cu")]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "host_cuda_shim.h"),
                   os.path.join(HERE, "emu_preprocess.cpp"), __file__, os.path.join(ROOT, "include", "scgr.h")]
    if os.path.exists(LIB_PRE) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PRE) for d in deps):
        return LIB_PRE
    os.makedirs(OUT_DIR, exist_ok=True)
    _common_host()
    body = open(srcs[0]).read().replace('#include "common.cuh"', "")
    with open(os.path.join(OUT_DIR, "preprocess_body.inc"), "w") as f:
        f.write(_rewrite_launches(body, 5))
    body = open(srcs[1]).read().replace('#include "common.cuh"', "")
    for pat, rep in _VOLATILE:
        body, n = re.subn(pat, rep, body)
        assert n == 1, pat
    assert "asm" not in body
    body, n = re.subn(r"extern __shared__ __align__\(16\) uint32_t (\w+)\[\];", r"uint32_t* \1 = reinterpret_cast<uint32_t*>(g_dyn_smem);", body)
    assert n == 1, "binning.cu: expected one dynamic shared-memory array"
    with open(os.path.join(OUT_DIR, "binning_body.inc"), "w") as f:
        f.write(_rewrite_launches(body, 2))
    body = open(srcs[2]).read().replace('#include "common.cuh"', "")
    for name, new in _RENDER_PTX.items():
        body = _replace_fn_body(body, name, new)
    assert "asm" not in body, "render.cu: an inline-PTX statement is not covered by _RENDER_PTX"
    # the variant switches are read once per process in the product; the tests flip them between launches
    body, n = re.subn(r"static const int (minb|tma|packed) = env_int", r"const int \1 = env_int", body)
    assert n == 5, n
    body = body.replace("int env_int(const char* name, int dflt) {", "int env_int_render(const char* name, int dflt) {").replace(
        "env_int(", "env_int_render(").replace("int env_int_render_render(", "int env_int_render(")
    with open(os.path.join(OUT_DIR, "render_body.inc"), "w") as f:
        f.write(_rewrite_launches(body, 0))
    _compile(LIB_PRE, "emu_preprocess.cpp")
    return LIB_PRE


def _common_host() -> None:
    c = open(os.path.join(CSRC, "common.cuh")).read().replace("#include <cuda_runtime.h>", "")
    c = c.replace('#include "../../include/scgr.h"', '#include "../../../include/scgr.h"')
    c, n = re.subn(r'asm\("sqrt\.approx\.ftz\.f32 %0, %1;"[^;]*;', "y = sqrtf(x);", c)
    assert n == 1, "common.cuh: expected exactly one sqrt.approx statement"
    for name, new in _TMA_PTX.items():
        c = _replace_fn_body(c, name, new)
    assert "asm" not in c, "common.cuh: an inline-PTX statement is not covered"
    with open(os.path.join(OUT_DIR, "common_host.cuh"), "w") as f:
        f.write(c)


LIB_LOSS = os.path.join(OUT_DIR, "libemu_loss_knn" + _SUFFIX)


def build_loss_knn() -> str:
    srcs = [os.path.join(CSRC, "loss.cu"), os.path.join(CSRC, "knn.cu"), os.path.join(CSRC, "prior.cu")]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "host_cuda_shim.h"),
                   os.path.join(HERE, "emu_loss_knn.cpp"), __file__, os.path.join(ROOT, "include", "scgr.h")]
    if os.path.exists(LIB_LOSS) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_LOSS) for d in deps):
        return LIB_LOSS
    os.makedirs(OUT_DIR, exist_ok=True)
    _common_host()
    for src, name, n in ((srcs[0], "loss_body.inc", 3), (srcs[1], "knn_body.inc", 1), (srcs[2], "prior_body.inc", 6)):
        body = open(src).read().replace('#include "common.cuh"', "")
        with open(os.path.join(OUT_DIR, name), "w") as f:
            f.write(_rewrite_launches(body, n))
    _compile(LIB_LOSS, "emu_loss_knn.cpp")
    return LIB_LOSS


LIB_FULL = os.path.join(OUT_DIR, "libemu_scgr" + _SUFFIX)


def build_full() -> str:
    """The WHOLE library on the host behind its real C entry points -- capi.cu included -- with every exported symbol
    renamed scgr_* -> emu_scgr_* so that nothing but a test can bind it (scgaussian_b200/_lib.py resolves scgr_* names
    only: this file cannot stand in for libscgr.so).  One translation unit per .cu file, as in the product build."""
    names = ["capi", "preprocess", "binning", "render", "loss", "knn", "model", "prior"]
    srcs = [os.path.join(CSRC, n + ".cu") for n in names]
    header = os.path.join(ROOT, "include", "scgr.h")
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "host_cuda_shim.h"), __file__, header]
    if os.path.exists(LIB_FULL) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_FULL) for d in deps):
        return LIB_FULL
    os.makedirs(OUT_DIR, exist_ok=True)
    build_preprocess()          # writes common_host.cuh and the preprocess / binning / render bodies
    build_loss_knn()
    build()
    hdr = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    symbols = sorted(set(re.findall(r"\b(scgr_[a-z0-9_]+)\s*\(", hdr)))
    with open(os.path.join(OUT_DIR, "rename_abi.h"), "w") as f:
        f.write("// generated: the emulated library exports emu_scgr_* only\n")
        f.writelines(f"#define {s} emu_{s}\n" for s in symbols)
    body = open(srcs[0]).read().replace('#include "common.cuh"', "")
    with open(os.path.join(OUT_DIR, "capi_body.inc"), "w") as f:
        f.write(body)
    gxx = shutil.which("g++")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    extra = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if ASAN else []
    objs = []
    for n in names:
        tu = os.path.join(OUT_DIR, f"tu_{n}.cpp")
        with open(tu, "w") as f:
            f.write('#define __CUDACC__ 1\n#include "rename_abi.h"\n#include "../host_cuda_shim.h"\n#include <stdexcept>\n#include <string>\n'
                    f'#include "common_host.cuh"\n#include "{n}_body.inc"\n')
            if n == "capi":     # the one launcher that is not emulated: NVSwitch multicast PTX
                f.write("namespace scgr { void launch_nvls_allreduce(void*, size_t, int, int, const Launch&) {\n"
                        '    throw std::runtime_error("scgr: the NVLS collective is not part of the host emulation"); }\n'
                        "void launch_nvls_allreduce_rows(void*, const float*, long long, int, int, int, const Launch&) {\n"
                        '    throw std::runtime_error("scgr: the NVLS collective is not part of the host emulation"); }\n'
                        "void launch_nvls_allreduce_fused(const ScgrNvlsFused&, const Launch&) {\n"
                        '    throw std::runtime_error("scgr: the NVLS collective is not part of the host emulation"); } }\n')
        obj = os.path.join(OUT_DIR, f"tu_{n}{'_asan' if ASAN else ''}.o")
        subprocess.check_call([gxx, "-O1", "-std=c++20", "-pthread", "-fPIC", "-ffp-contract=off", "-w", *extra, "-c", "-o", obj, tu],
                              env=env, cwd=OUT_DIR)
        objs.append(obj)
    subprocess.check_call([gxx, "-shared", "-pthread", *extra, "-o", LIB_FULL, *objs], env=env, cwd=OUT_DIR)
    return LIB_FULL


if __name__ == "__main__":
    print(build(), build_preprocess(), build_loss_knn(), build_full())
