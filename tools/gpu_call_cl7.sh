#!/bin/bash
# smem-bin summation tree A/B, cluster tests, plan timing (256 vs 512 extraction threads)
mkdir -p gpurun_out; rm -f gpurun_out/r02cl7_rc.txt
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
  timeout 300 python tools/phase_timers.py > gpurun_out/r02cl7_phases_$n.log 2>&1; echo "$n phases rc=$?" >> gpurun_out/r02cl7_rc.txt
  grep -E "nodes \[(3|0)\]" gpurun_out/r02cl7_phases_$n.log | cut -c1-260
  GNNX_STUDY_TAG=_ab7_${n} timeout 300 python tools/cluster_study.py syn1 0 1 > gpurun_out/r02cl7_ab_${n}.log 2>&1; echo "$n rc=$?" >> gpurun_out/r02cl7_rc.txt
  show gpurun_out/cluster_study_syn1_ab7_${n}.json
done
unset GNNX_LIB_PATH
timeout 600 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_state.py -x -q > gpurun_out/r02cl7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02cl7_rc.txt; tail -3 gpurun_out/r02cl7_tests.log
for n in default kh512; do
  unset GNNX_LIB_PATH
  if [ $n != default ]; then export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_$n/libgnnx.so; fi
  echo "== plan $n"; timeout 200 python tools/plan_time.py 2>gpurun_out/r02cl7_plan_$n.err | tee gpurun_out/r02cl7_plan_$n.json
done
unset GNNX_LIB_PATH
GNNX_HOST_TIMING=1 timeout 200 python tools/plan_time.py 2>&1 | grep "gnnx\]" | tail -4
cat gpurun_out/r02cl7_rc.txt
