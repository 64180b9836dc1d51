#!/bin/bash
# where does the fixed summation tree cost time?  phase timers + full-batch A/B of three builds
mkdir -p gpurun_out; rm -f gpurun_out/r02cl5_rc.txt
show() { python - "$1" <<'P'
import json, sys
for r in json.load(open(sys.argv[1])):
    print(r['batch'], r['cluster'], round(r['kernel_ms_min'],3), round(r['kernel_ms_med'],3), r['class_counts'], 'end', r['class_end_ms'])
P
}
for n in default vwoff default vwoff; do
  unset GNNX_LIB_PATH
  if [ $n != default ]; then export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_$n/libgnnx.so; fi
  echo "== $n"
  timeout 300 python tools/phase_timers.py > gpurun_out/r02cl5_phases_$n.log 2>&1; echo "$n phases rc=$?" >> gpurun_out/r02cl5_rc.txt
  cut -c1-260 gpurun_out/r02cl5_phases_$n.log
  GNNX_STUDY_TAG=_ab5_${n} timeout 300 python tools/cluster_study.py syn1 0 1 > gpurun_out/r02cl5_ab_${n}.log 2>&1; echo "$n rc=$?" >> gpurun_out/r02cl5_rc.txt
  show gpurun_out/cluster_study_syn1_ab5_${n}.json
done
cat gpurun_out/r02cl5_rc.txt
