#!/bin/bash
# round-2 GPU call C: test suite (optimiser variants, pipelined gang kernel), gang-size throughput study at full concurrency
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02c_rc.txt
timeout 900 python tools/gang_study.py 100000 148 4 -1,8,18,37,74,148 > gpurun_out/r02c_gang100k_full.log 2>&1; echo "gang100k rc=$?" >> gpurun_out/r02c_rc.txt
timeout 600 python tools/gang_study.py 20000 148 6 -1,1,4,8,18,37 > gpurun_out/r02c_gang20k_full.log 2>&1; echo "gang20k rc=$?" >> gpurun_out/r02c_rc.txt
timeout 300 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02c_gang100k_4.log 2>&1; echo "gang100k_4 rc=$?" >> gpurun_out/r02c_rc.txt
tail -n 25 gpurun_out/r02c_pytest.log | cut -c1-300; cut -c1-500 gpurun_out/r02c_gang100k_full.log; cut -c1-500 gpurun_out/r02c_gang20k_full.log; cut -c1-500 gpurun_out/r02c_gang100k_4.log; cat gpurun_out/r02c_rc.txt
