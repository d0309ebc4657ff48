#!/bin/bash
# Round-2 GPU pass B (1 GPU): A/B of the packed-fp32 render variants, then the parity suite (strict) on the default.
set -u
mkdir -p gpurun_out
python tools/ab.py "SCGR_FWD_PACKED=0 SCGR_BWD_MINB=18" "SCGR_FWD_PACKED=0 SCGR_BWD_MINB=16" "SCGR_FWD_PACKED=0 SCGR_BWD_MINB=20" \
   "SCGR_FWD_PACKED=1 SCGR_FWD_MINB=20 SCGR_BWD_MINB=16" "SCGR_FWD_PACKED=1 SCGR_FWD_MINB=18 SCGR_BWD_MINB=16" \
   "SCGR_FWD_PACKED=1 SCGR_FWD_MINB=16 SCGR_BWD_MINB=16" "SCGR_FWD_PACKED=1 SCGR_FWD_MINB=16 SCGR_BWD_MINB=14" \
   "SCGR_FWD_PACKED=1 SCGR_FWD_MINB=16 SCGR_BWD_MINB=16 SCGR_TMA_BWD=0" > gpurun_out/b_ab.log 2>&1
cat gpurun_out/b_ab.log
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
