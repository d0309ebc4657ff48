#!/bin/bash
# Round-2 GPU pass T (1 GPU): the two switches written without a GPU -- SCGR_PDL=1 (programmatic dependent launch) and
# SCGR_LOSS_VARIANT=1 (streaming loss kernels).  Whole -m gpu suite under SCGR_PDL=1 (the loss tests run both variants
# by themselves), loss kernels alone, A/B of the bench value (PDL) and of the end-to-end entry (both switches).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
SCGR_PDL=1 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_pytest_pdl.log 2>&1
echo "pytest(SCGR_PDL=1) rc=$?" | tee -a gpurun_out/t_pytest_pdl.log; tail -4 gpurun_out/t_pytest_pdl.log | cut -c1-300
cp gpurun_out/parity_report.jsonl gpurun_out/t_parity_pdl.jsonl 2>/dev/null
timeout 300 python tools/loss_only.py 20 > gpurun_out/t_loss_only.log 2>&1; cat gpurun_out/t_loss_only.log | cut -c1-260
FAST="--no-cpu-baseline --no-train-step --no-standin --no-batch8 --no-config2"
for rep in 1 2; do for pdl in 0 1; do
  SCGR_PDL=$pdl timeout 300 python bench.py --steps 40 --warmup 8 $FAST --no-e2e > gpurun_out/t_value_pdl${pdl}_$rep.json 2> gpurun_out/t_value_pdl${pdl}_$rep.err || tail -c 400 gpurun_out/t_value_pdl${pdl}_$rep.err
done; done
for cfg in "0 0" "1 0" "0 1" "1 1"; do set -- $cfg
  SCGR_PDL=$1 SCGR_LOSS_VARIANT=$2 timeout 300 python bench.py --steps 24 --warmup 8 $FAST > gpurun_out/t_e2e_pdl$1_loss$2.json 2> gpurun_out/t_e2e_pdl$1_loss$2.err || tail -c 400 gpurun_out/t_e2e_pdl$1_loss$2.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/t_value_*.json") + glob.glob("gpurun_out/t_e2e_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    e2e = d.get("e2e") or {}
    print(f.split("/")[-1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(e2e.get("value", 0), 1),
          "graph", round((e2e.get("graph_replay") or {}).get("value", 0), 1),
          json.dumps({k: round(x["ms_per_step"] * 1000, 1) for k, x in (d.get("kernels") or {}).items()}))
PY
