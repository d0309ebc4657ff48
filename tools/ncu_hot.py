"""Where a kernel's warps wait: reads `ncu -i REP --page source --csv --print-source sass` of one launch and prints
(a) the stall samples and executed instructions of every PHASE of the kernel (a phase = the SASS between two block
barriers, in address order), (b) the hottest instructions.  Development tool (profiles/ summaries are written from it).

    python tools/ncu_hot.py REP.ncu-rep [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:150])
    h = rows[1]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [(k, i) for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
    body = [r for r in rows[2:] if len(r) > iE and (r[iN] or '0').isdigit() and (r[iE] or '0').isdigit()]
    tot = sum(int(r[iN] or 0) for r in body) or 1
    tot_i = sum(int(r[iE] or 0) for r in body) or 1
    print(f"samples {tot}, warp instructions {tot_i}, static instructions {len(body)}")
    phase, acc, acc_i, n0 = 0, 0, 0, 0
    reasons = {}
    print("phase  first..last(static idx)  samples%  instr%  top stall reasons")
    for n, r in enumerate(body):
        s = int(r[iN] or 0)
        acc += s
        acc_i += int(r[iE] or 0)
        for k, i in stalls:
            v = int(r[i] or 0)
            if v:
                reasons[k] = reasons.get(k, 0) + v
        if "BAR.SYNC" in r[iS] or n == len(body) - 1:
            rs = sorted(reasons.items(), key=lambda kv: -kv[1])[:3]
            print(f"{phase:3d}   {n0:5d}..{n:5d}   {100 * acc / tot:6.1f}  {100 * acc_i / tot_i:6.1f}   " +
                  ", ".join(f"{k[6:]} {100 * v / tot:.1f}" for k, v in rs))
            phase, acc, acc_i, n0, reasons = phase + 1, 0, 0, n + 1, {}
    print("hottest instructions:")
    order = sorted(range(len(body)), key=lambda n: -int(body[n][iN] or 0))[:top]
    for n in sorted(order):
        r = body[n]
        rs = sorted(((k, int(r[i] or 0)) for k, i in stalls), key=lambda kv: -kv[1])[:2]
        print(f"  [{n:5d}] {100 * int(r[iN] or 0) / tot:5.1f}%  x{int(r[iE] or 0):<8d} {r[iS].strip()[:70]:70s} " +
              ", ".join(f"{k[6:]} {v}" for k, v in rs if v))


if __name__ == "__main__":
    main()
