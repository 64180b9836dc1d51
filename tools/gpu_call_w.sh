#!/bin/bash
# round-2 GPU call W: wide models (hidden / output width 33..128) through the variant kernel; full GPU suite; sanitizers on the variant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02w_rc.txt
SAN_EPOCHS=3 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py var > gpurun_out/r02w_mem_var.log 2>&1; echo "mem var rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r02w_mem_var.log | tail -1)" >> gpurun_out/r02w_rc.txt
SAN_EPOCHS=3 timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/sanitize_run.py var > gpurun_out/r02w_race_var.log 2>&1; echo "race var rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02w_race_var.log | tail -1)" >> gpurun_out/r02w_rc.txt
cat gpurun_out/r02w_rc.txt; tail -n 15 gpurun_out/r02w_pytest.log | cut -c1-300
