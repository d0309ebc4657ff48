"""A/B helper (GPU box): runs bench.py under several environment settings and prints the per-kernel
milliseconds side by side.  usage: python tools/ab.py "SCGR_FWD_WPT=1 SCGR_BWD_WPT=1" "SCGR_FWD_WPT=2" ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = sys.argv[1:] or [""]
rows = {}
for v in variants:
    env = dict(os.environ)
    for kv in v.split():
        k, val = kv.split("=")
        env[k] = val
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3",
                          "--no-e2e", "--no-cpu-baseline", "--no-train-step", "--no-standin", "--no-batch8"],
                         env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print("FAILED", v, out.stderr[-2000:])
        continue
    d = json.loads(line[-1])
    rows[v] = d
    print(f"== [{v}] {d['value']:.1f} views/s  {d['ms_per_step']:.3f} ms/step  R={d['config']['num_rendered_R_mean']:.0f}")
    print("   " + "  ".join(f"{k}={x['ms_per_step']*1000:.0f}us" for k, x in d["kernels"].items()))
