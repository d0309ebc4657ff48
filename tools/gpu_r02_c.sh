#!/bin/bash
# Round-2 GPU pass C (1 GPU): strict parity suite; ncu --set full of render_backward, scalar (default) and packed.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -8 gpurun_out/c_pytest.log
BFLAGS="--steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --no-standin --no-batch8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_backward -s 5 -c 1 -f -o gpurun_out/c_bwd_scalar python bench.py $BFLAGS > gpurun_out/c_ncu_scalar.log 2>&1
SCGR_BWD_PACKED=1 SCGR_BWD_MINB=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_backward -s 5 -c 1 -f -o gpurun_out/c_bwd_packed python bench.py $BFLAGS > gpurun_out/c_ncu_packed.log 2>&1
ls -la gpurun_out/*.ncu-rep
