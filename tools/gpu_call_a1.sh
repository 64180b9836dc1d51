#!/bin/bash
# round-2 GPU call A1: cluster sanity, GPU test suite (all failures), cluster study, parity report, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 300 python tools/cluster_study.py syn1 > gpurun_out/r02a_cluster_syn1.log 2>&1; echo "cluster study rc=$?" > gpurun_out/r02a_rc.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_rc.txt
timeout 400 python tools/parity_report.py > gpurun_out/r02a_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02a_rc.txt
cp gpurun_out/parity_report.json gpurun_out/r02a_parity_report.json
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_rc.txt
tail -n 30 gpurun_out/r02a_pytest.log; tail -n 12 gpurun_out/r02a_cluster_syn1.log; tail -c 600 gpurun_out/r02a_parity.log; cat gpurun_out/r02a_rc.txt; tail -c 1500 gpurun_out/r02a_bench.json; tail -c 800 gpurun_out/r02a_bench.err
