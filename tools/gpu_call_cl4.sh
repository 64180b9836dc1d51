#!/bin/bash
# carveout + fixed-tree reduction: cluster auto study, A/B on the full batch
mkdir -p gpurun_out; rm -f gpurun_out/r02cl4_rc.txt
show() { python - "$1" <<'P'
import json, sys
for r in json.load(open(sys.argv[1])):
    print(r['batch'], r['cluster'], round(r['kernel_ms_min'],3), round(r['kernel_ms_med'],3), r['class_counts'], r['plan_cluster'], r['checksum_equals_off'], 'end', r['class_end_ms'])
P
}
GNNX_STUDY_TAG=_auto4 timeout 600 python tools/cluster_study.py syn1 1,24,88,130,175,260,350,0 1,0 > gpurun_out/r02cl4_auto.log 2>&1; echo "auto rc=$?" >> gpurun_out/r02cl4_rc.txt
show gpurun_out/cluster_study_syn1_auto4.json
for rep in 1 2; do
for n in default nocarve vwoff; do
  unset GNNX_LIB_PATH GNNX_CARVEOUT
  if [ $n = vwoff ]; then export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_vwoff/libgnnx.so; fi
  if [ $n = nocarve ]; then export GNNX_CARVEOUT=0; fi
  GNNX_STUDY_TAG=_ab4_${n}_$rep timeout 300 python tools/cluster_study.py syn1 0 1 > gpurun_out/r02cl4_ab_${n}_$rep.log 2>&1; echo "$n $rep rc=$?" >> gpurun_out/r02cl4_rc.txt
  echo $n $rep; show gpurun_out/cluster_study_syn1_ab4_${n}_$rep.json
done
done
unset GNNX_LIB_PATH GNNX_CARVEOUT
timeout 300 python bench.py --steps 10 --warmup 3 --workload syn4 --no-cpu > gpurun_out/r02cl4_syn4.json 2>gpurun_out/r02cl4_syn4.err; cut -c1-330 gpurun_out/r02cl4_syn4.json
GNNX_CARVEOUT=0 timeout 300 python bench.py --steps 10 --warmup 3 --workload syn4 --no-cpu > gpurun_out/r02cl4_syn4_nocarve.json 2>gpurun_out/r02cl4_syn4.err; cut -c1-330 gpurun_out/r02cl4_syn4_nocarve.json
cat gpurun_out/r02cl4_rc.txt
