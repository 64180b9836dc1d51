#!/bin/bash
# automatic cluster class: tests, study, dist overhead
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_parity.py tests/test_gpu_state.py tests/test_gpu_dist.py -x -q > gpurun_out/r02cl_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02cl_rc.txt
tail -5 gpurun_out/r02cl_tests.log
timeout 600 python tools/cluster_study.py syn1 > gpurun_out/r02cl_study.log 2>&1; echo "study rc=$?" >> gpurun_out/r02cl_rc.txt
cut -c1-600 gpurun_out/r02cl_study.log
timeout 300 python tools/dist_overhead.py 8 > gpurun_out/r02cl_dist.log 2>&1; echo "dist rc=$?" >> gpurun_out/r02cl_rc.txt
cut -c1-250 gpurun_out/r02cl_dist.log | head -70
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/r02cl_bench.json 2> gpurun_out/r02cl_bench.err; echo "bench rc=$?" >> gpurun_out/r02cl_rc.txt
cut -c1-400 gpurun_out/r02cl_bench.json
cat gpurun_out/r02cl_rc.txt
