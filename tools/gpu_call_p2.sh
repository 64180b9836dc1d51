#!/bin/bash
# plan timing: sub-warp row groups (8 lanes per induced row) in the k-hop kernels vs one row per warp
mkdir -p gpurun_out; rm -f gpurun_out/r02p2_rc.txt
for n in default sub default sub; do
  unset GNNX_LIB_PATH
  if [ $n != default ]; then export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_$n/libgnnx.so; fi
  echo "== plan $n"; timeout 200 python tools/plan_time.py 2>gpurun_out/r02p2_plan_$n.err | tee gpurun_out/r02p2_plan_$n.json
done
export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_sub/libgnnx.so
GNNX_HOST_TIMING=1 timeout 200 python tools/plan_time.py 2>&1 | grep "gnnx\]" | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py tests/test_gpu_graph_mode.py -x -q > gpurun_out/r02p2_tests_sub.log 2>&1; echo "tests sub rc=$?" >> gpurun_out/r02p2_rc.txt; tail -2 gpurun_out/r02p2_tests_sub.log
timeout 300 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --c5-n 20000 > gpurun_out/r02p2_c5_20k_sub.json 2> gpurun_out/r02p2_c5.err; echo "c5 20k sub rc=$?" >> gpurun_out/r02p2_rc.txt; cut -c1-300 gpurun_out/r02p2_c5_20k_sub.json
cat gpurun_out/r02p2_rc.txt
