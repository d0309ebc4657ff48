#!/bin/bash
# Round-2 GPU pass A (1 GPU): the -m gpu suite in calibration mode (records instead of failing on unexplained parity
# elements, so that one run sizes the error model), the bench line, the ncu launch list and the instruction counts of
# the render kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
SCGR_PARITY_CALIBRATE=1 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -40 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/a_bench.err; head -c 600 gpurun_out/a_bench.json
BFLAGS="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --no-standin --no-batch8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/a_launches.csv python bench.py $BFLAGS > gpurun_out/a_ncu1.log 2>&1
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:render_ -s 12 -c 32 --csv --log-file gpurun_out/a_issue.csv python bench.py $BFLAGS > gpurun_out/a_ncu2.log 2>&1
echo done
