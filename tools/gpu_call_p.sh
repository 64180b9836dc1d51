#!/bin/bash
# plan timing: host classification under the fill kernel; 512 vs 1024 extraction threads
mkdir -p gpurun_out; rm -f gpurun_out/r02p_rc.txt
for n in default kh1024 default kh1024; do
  unset GNNX_LIB_PATH
  if [ $n != default ]; then export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_$n/libgnnx.so; fi
  echo "== plan $n"; timeout 200 python tools/plan_time.py 2>gpurun_out/r02p_plan_$n.err | tee gpurun_out/r02p_plan_$n.json
done
unset GNNX_LIB_PATH
GNNX_HOST_TIMING=1 timeout 200 python tools/plan_time.py 2>&1 | grep "gnnx\]" | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py -x -q > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p_rc.txt; tail -2 gpurun_out/r02p_tests.log
GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_kh1024/libgnnx.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "khop or neighborhood or extract or dropin" > gpurun_out/r02p_tests_kh1024.log 2>&1; echo "tests kh1024 rc=$?" >> gpurun_out/r02p_rc.txt; tail -2 gpurun_out/r02p_tests_kh1024.log
cat gpurun_out/r02p_rc.txt
