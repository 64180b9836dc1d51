#!/bin/bash
# round-2 GPU call B: test suite with the gang kernel + resume fix, gang study (N=20k), cluster small-batch study, c5 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02b_rc.txt
timeout 600 python tools/gang_study.py 20000 16 12 -1,0,1,8,37,74,148 > gpurun_out/r02b_gang20k.log 2>&1; echo "gang20k rc=$?" >> gpurun_out/r02b_rc.txt
timeout 600 python tools/gang_study.py 100000 4 6 -1,0,37,148 > gpurun_out/r02b_gang100k.log 2>&1; echo "gang100k rc=$?" >> gpurun_out/r02b_rc.txt
timeout 400 python tools/cluster_study.py syn1 > gpurun_out/r02b_cluster_syn1.log 2>&1; echo "cluster rc=$?" >> gpurun_out/r02b_rc.txt
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/r02b_rc.txt
tail -n 25 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_gang20k.log | cut -c1-400; cat gpurun_out/r02b_gang100k.log | cut -c1-400; tail -n 14 gpurun_out/r02b_cluster_syn1.log | cut -c1-300; cat gpurun_out/r02b_rc.txt; tail -c 1500 gpurun_out/r02b_bench_c5.json; tail -c 600 gpurun_out/r02b_bench_c5.err
