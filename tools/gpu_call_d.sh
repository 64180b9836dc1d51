#!/bin/bash
# round-2 GPU call D: sanitizers (racecheck + memcheck) on every kernel family of the final build
mkdir -p gpurun_out
: > gpurun_out/r02d_rc.txt
for part in node stream graph misc var cluster; do
  SAN_EPOCHS=3 timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/sanitize_run.py $part > gpurun_out/r02d_race_$part.log 2>&1
  echo "race $part rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02d_race_$part.log | tail -1)" >> gpurun_out/r02d_rc.txt
done
for part in node stream graph misc var cluster; do
  SAN_EPOCHS=3 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py $part > gpurun_out/r02d_mem_$part.log 2>&1
  echo "mem $part rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r02d_mem_$part.log | tail -1)" >> gpurun_out/r02d_rc.txt
done
cat gpurun_out/r02d_rc.txt
