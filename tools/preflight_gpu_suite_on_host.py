"""Pre-flight of the `-m gpu` tests WITHOUT a GPU (development tool; GPU minutes are the scarce resource).

What it does: copies the repository into a scratch directory, REWRITES THE COPY -- the "CUDA tensors only" checks, the
stream / event / pinned-memory calls, the `"cuda"` device strings of the tests -- and binds the copy's `_lib.load()` to
the host-emulated library of tests/emulation (libemu_scgr.so: the real kernel sources and the real capi.cu compiled with
g++, exported as emu_scgr_*).  Then it runs the copy's GPU tests on CPU tensors.  The product's own Python host code
(rasterizer.py, losses.py, model.py, optim.py, densify.py) therefore drives the real C entry points and the real kernel
logic, thread for thread -- slowly.

What it is for: finding Python-level mistakes in new GPU tests, wrong pointer / struct plumbing, kernel logic errors and
cross-test state problems (it runs the files in the order `pytest -x` will) BEFORE the suite is sent to a B200.
What it is not: a CPU path of the product.  Nothing in the repository is modified, nothing under scgaussian_b200/ can
load the emulated library (it exports no scgr_* symbol), and no result of this tool is a parity or performance claim --
the B200 run is.  Full-size cases (1M Gaussians, 4K, 1080p loss) are deselected: the emulation runs one OS thread per
CUDA thread.

    python tools/preflight_gpu_suite_on_host.py [extra pytest args]
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SKIP = os.environ.get("PREFLIGHT_SKIP", "not config3 and not config4 and not full_size and not 70001 and not 378")


def main() -> int:
    from tests.emulation import build
    emu = build.build_full()
    work = tempfile.mkdtemp(prefix="scgr_preflight_")
    for d in ("scgaussian_b200", "oracle", "tests", "include", "diff_gaussian_rasterization", "simple_knn", "tools"):
        shutil.copytree(os.path.join(ROOT, d), os.path.join(work, d),
                        ignore=shutil.ignore_patterns("__pycache__", "_build", "*.so", "*.o"))
    for so in glob.glob(os.path.join(ROOT, "oracle", "*.so")):
        shutil.copy(so, os.path.join(work, "oracle"))
    # ---- the copy of the host code: no device checks, no CUDA runtime calls
    for f in glob.glob(os.path.join(work, "scgaussian_b200", "*.py")) + [os.path.join(work, "simple_knn", "_C.py"),
                                                                         os.path.join(work, "tools", "train_step_time.py")]:
        s = open(f).read()
        s = s.replace('.type != "cuda"', '.type == "never"').replace(".pin_memory()", "")
        s = re.sub(r"torch\.cuda\.device\(([^)]*)\)", "__import__('contextlib').nullcontext()", s)
        s = re.sub(r"torch\.cuda\.current_stream\(([^)]*)\)\.cuda_stream", "0", s)
        s = re.sub(r"torch\.cuda\.current_stream\(([^)]*)\)\.synchronize\(\)", "None", s)
        s = s.replace("torch.cuda.current_device()", "0").replace("torch.cuda.synchronize()", "None")
        open(f, "w").write(s)
    p = os.path.join(work, "scgaussian_b200", "_lib.py")
    s = open(p).read()
    old = '''    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the export is missing'''
    assert old in s
    s = s.replace(old, f'''    raw = C.CDLL({emu!r})
    lib = type("EmulatedLibrary", (), {{}})()
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(raw, "emu_" + name)
        setattr(lib, name, fn)''')
    s = s.replace("if not os.path.exists(LIB_PATH):", "if False:")
    open(p, "w").write(s)
    # scratch buffers: the C ABI wants 256-byte alignment, which the CPU allocator does not promise
    p = os.path.join(work, "scgaussian_b200", "rasterizer.py")
    s = open(p).read()
    old = "    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)"
    assert old in s
    s = s.replace(old, "    buf = torch.empty(max(int(nbytes), 256) + 256, dtype=torch.uint8, device=device)\n"
                       "    off = (-buf.data_ptr()) % 256\n    return buf[off:off + max(int(nbytes), 256)]")
    open(p, "w").write(s)
    p = os.path.join(work, "scgaussian_b200", "losses.py")
    s = open(p).read()
    old = "scratch = torch.empty(lib.scgr_photometric_scratch_bytes(c, h, w), dtype=torch.uint8, device=dev)"
    assert old in s
    s = s.replace(old, "_b = torch.empty(lib.scgr_photometric_scratch_bytes(c, h, w) + 256, dtype=torch.uint8, device=dev)\n"
                       "            scratch = _b[(-_b.data_ptr()) % 256:]")
    open(p, "w").write(s)
    # ---- the copy of the tests: CPU tensors
    for f in glob.glob(os.path.join(work, "tests", "*.py")):
        s = open(f).read()
        s = s.replace("torch.cuda.is_available()", "True")
        s = s.replace('torch.device("cuda:0")', 'torch.device("cpu")').replace('torch.device("cuda", 0)', 'torch.device("cpu")')
        s = s.replace('device="cuda"', 'device="cpu"').replace('dev = "cuda"', 'dev = "cpu"').replace(".cuda()", "")
        s = s.replace("torch.cuda.synchronize()", "None")
        open(f, "w").write(s)
    # the staged copy of the reference's Python files (oracle/_ref/reference, made by __graft_entry__.build()): CPU too
    for f in glob.glob(os.path.join(work, "oracle", "_ref", "reference", "**", "*.py"), recursive=True):
        s = open(f).read()
        open(f, "w").write(s.replace('device="cuda"', 'device="cpu"').replace("device='cuda'", "device='cpu'").replace(".cuda()", ""))
    os.environ["SCGR_REFERENCE_DIR"] = os.path.join(work, "oracle", "_ref", "reference")
    p = os.path.join(work, "tests", "conftest.py")
    s = open(p).read().replace("        b.build_library()", "        pass      # pre-flight: the emulated library stands in")
    open(p, "w").write(s)
    cmd = [sys.executable, "-m", "pytest", "tests", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-k", SKIP, *sys.argv[1:]]
    print("pre-flight in", work, "\n ", " ".join(cmd), flush=True)
    rc = subprocess.call(cmd, cwd=work)
    shutil.rmtree(work, ignore_errors=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
