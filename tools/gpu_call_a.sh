#!/bin/bash
# round-2 GPU call A: GPU test suite, parity report, sanitizers, baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02a_san_rc.txt
timeout 600 python tools/parity_report.py > gpurun_out/r02a_parity.log 2>&1
cp gpurun_out/parity_report.json gpurun_out/r02a_parity_report.json
for part in node stream graph misc; do
  SAN_EPOCHS=3 timeout 420 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 30 python tools/sanitize_run.py $part > gpurun_out/r02a_race_$part.log 2>&1
  echo "race $part rc=$?" >> gpurun_out/r02a_san_rc.txt
done
for part in node stream graph misc; do
  SAN_EPOCHS=3 timeout 300 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize_run.py $part > gpurun_out/r02a_mem_$part.log 2>&1
  echo "mem $part rc=$?" >> gpurun_out/r02a_san_rc.txt
done
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -n 15 gpurun_out/r02a_pytest.log; tail -c 600 gpurun_out/r02a_parity.log; cat gpurun_out/r02a_san_rc.txt; tail -c 1200 gpurun_out/r02a_bench.json
