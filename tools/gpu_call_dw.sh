#!/bin/bash
# A/B: dense-warp count / tile buffering of the gang kernel's tensor-core passes
mkdir -p gpurun_out
for n in default dw7 dw12b1 dw14b1; do
  if [ $n = default ]; then unset GNNX_LIB_PATH; else export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_$n/libgnnx.so; fi
  timeout 300 python tools/gang_study.py 100000 4 6 0 > gpurun_out/r02dw_$n.log 2>&1; echo "$n rc=$?" >> gpurun_out/r02dw_rc.txt
  echo $n; cut -c1-420 gpurun_out/r02dw_$n.log
done
cat gpurun_out/r02dw_rc.txt
