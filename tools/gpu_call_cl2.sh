#!/bin/bash
# automatic cluster class: timeline of the launch classes, stream priorities
mkdir -p gpurun_out
for p in 0 1 2; do
  GNNX_STREAM_PRIO=$p timeout 400 python tools/cluster_study.py syn1 88,130,175,260,0 1,0 > gpurun_out/r02cl2_prio$p.log 2>&1; echo "prio$p rc=$?" >> gpurun_out/r02cl2_rc.txt
  cut -c1-120 gpurun_out/r02cl2_prio$p.log; grep -o '"class_counts.*' gpurun_out/r02cl2_prio$p.log | cut -c1-420
done
cat gpurun_out/r02cl2_rc.txt
