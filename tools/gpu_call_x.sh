#!/bin/bash
# round-2 GPU call X: graph-mode footprint classes -- graph tests, graphs bench, sanitizer on graph mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph_mode.py tests/test_gpu_state.py -q --timeout 300 > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02x_rc.txt
timeout 600 python bench.py --workload graphs --steps 5 --warmup 3 > gpurun_out/r02x_bench_graphs.json 2> gpurun_out/r02x_bench_graphs.err; echo "bench graphs rc=$?" >> gpurun_out/r02x_rc.txt
SAN_EPOCHS=3 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py graph > gpurun_out/r02x_mem_graph.log 2>&1; echo "mem graph rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r02x_mem_graph.log | tail -1)" >> gpurun_out/r02x_rc.txt
cat gpurun_out/r02x_rc.txt; tail -n 5 gpurun_out/r02x_pytest.log | cut -c1-300; python -c "
import json
d=json.loads(open('gpurun_out/r02x_bench_graphs.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
"; tail -c 300 gpurun_out/r02x_bench_graphs.err
