#!/bin/bash
# automatic cluster class with whole-SM cluster CTAs; A/B of the fixed summation tree (vwarps) on the full batch
mkdir -p gpurun_out
GNNX_STUDY_TAG=_auto timeout 600 python tools/cluster_study.py syn1 1,24,88,130,175,260,350,0 1,0 > gpurun_out/r02cl3_auto.log 2>&1; echo "auto rc=$?" > gpurun_out/r02cl3_rc.txt
python - <<'P'
import json
for r in json.load(open('gpurun_out/cluster_study_syn1_auto.json')):
    print(r['batch'], r['cluster'], round(r['kernel_ms_min'],3), r['class_counts'], r['plan_cluster'], r['checksum_equals_off'], 'end', r['class_end_ms'])
P
for rep in 1 2; do
for n in default vwoff; do
  if [ $n = default ]; then unset GNNX_LIB_PATH; else export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_$n/libgnnx.so; fi
  GNNX_STUDY_TAG=_ab_${n}_$rep timeout 300 python tools/cluster_study.py syn1 0 1 > gpurun_out/r02cl3_ab_${n}_$rep.log 2>&1; echo "$n $rep rc=$?" >> gpurun_out/r02cl3_rc.txt
  echo $n $rep; grep -o '"kernel_ms_min": [0-9.]*, "kernel_ms_med": [0-9.]*' gpurun_out/r02cl3_ab_${n}_$rep.log; grep -o '"class_counts.*' gpurun_out/r02cl3_ab_${n}_$rep.log | cut -c1-300
done
done
unset GNNX_LIB_PATH
cat gpurun_out/r02cl3_rc.txt
