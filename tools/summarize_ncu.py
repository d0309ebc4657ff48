"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under
profiles/:  python tools/summarize_ncu.py <round-tag>
  gpurun_out/launches.csv   (ncu --metrics gpu__time_duration.sum ... bench.py)  -> profiles/<tag>_launches.md
  gpurun_out/prof_all.ncu-rep (ncu --set full ... one step)                       -> profiles/<tag>_kernels.md / traffic.json
A second argument names another capture (e.g. gpurun_out/prof_model.ncu-rep of tools/model_ncu.py): its kernels go to
profiles/<tag>_kernels.md and are merged into traffic.json.
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
ours = ("scgr::",)


def short(name):
    n = name.split("(")[0].replace("void ", "")
    for pre in ("scgr::<unnamed>::", "scgr::(anonymous namespace)::", "<unnamed>::", "unnamed>::", "scgr::"):
        n = n.replace(pre, "")
    return n


# ---------------- launch list ----------------
lp = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = OrderedDict()
    tot_ours = 0.0
    tot_other = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        k = short(r["Kernel Name"])
        is_ours = "scgr::" in r["Kernel Name"]
        d = per.setdefault(k, dict(n=0, ns=0.0, ours=is_ours, grid=r["Grid Size"], block=r["Block Size"]))
        d["n"] += 1
        d["ns"] += ns
        if is_ours:
            tot_ours += ns
        else:
            tot_other += ns
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag})\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py "
                "--steps 2 --warmup 3 --no-cpu-baseline --no-e2e` (7 steps of config 3 incl. the profiled loop; "
                "per-launch times are cold-cache and serialised: compare SHARES, not absolutes).\n\n")
        f.write("| kernel | launches | total us | avg us | share of libscgr time | grid | block |\n|---|---|---|---|---|---|---|\n")
        for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
            share = f"{100 * d['ns'] / tot_ours:.1f} %" if d["ours"] else "(torch)"
            f.write(f"| `{k}` | {d['n']} | {d['ns'] / 1e3:.1f} | {d['ns'] / 1e3 / d['n']:.1f} | {share} | {d['grid']} | {d['block']} |\n")
        f.write(f"\nlibscgr kernels: {tot_ours / 1e3:.1f} us total; other (torch fill/copy/elementwise): {tot_other / 1e3:.1f} us.\n")
    print("wrote", f"{tag}_launches.md")

# ---------------- full-set capture ----------------
rp = os.path.join(ROOT, sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "prof_all.ncu-rep")
what = "One step of config 3 (1M Gaussians, 1920x1080, SH3)" if len(sys.argv) <= 2 else \
    f"`{os.path.basename(rp)}` (1M Gaussians, SH3)"
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp insts"),
            ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst")]
    stall = [h for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h]
    traffic = {}
    with open(os.path.join(out_dir, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# ncu --set full summary ({tag})\n\n{what} under "
                "`ncu --set full --clock-control none --import-source on`.  Numbers under the profiler are NOT bench values.\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            k = short(d["Kernel Name"])
            f.write(f"## `{k}`  grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}\n\n")
            for m, label in want:
                if m in d:
                    f.write(f"- {label}: {d[m]} {units[hdr.index(m)]}\n")
            st = sorted(((float(d[h].replace(',', '')), h.split("stalled_")[1].split("_per")[0]) for h in stall), reverse=True)[:5]
            f.write("- top stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in st) + "\n\n")
            def mb(x, u):
                v = float(x.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            try:
                t = mb(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")]) + \
                    mb(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
                base = k.split("<")[0].replace("_kernel", "")
                traffic.setdefault(base, []).append(t)
            except Exception:
                pass
    tp = os.path.join(out_dir, "traffic.json")
    merged = json.load(open(tp)) if len(sys.argv) > 2 and os.path.exists(tp) else {}
    merged.update({k: sum(v) / len(v) for k, v in traffic.items()})
    json.dump(merged, open(tp, "w"), indent=1)
    print("wrote", f"{tag}_kernels.md", "traffic.json")
