#!/bin/bash
# round-2 GPU call Z: FINAL build -- full GPU tests, smoke, default bench + reference arm, contract launch list, memcheck of the node part
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02z_rc.txt
python __graft_entry__.py smoke > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02z_rc.txt
timeout 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?" >> gpurun_out/r02z_rc.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/r02z_rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02z_launches_syn1.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/r02z_launchrun.log 2>&1; echo "launch list rc=$?" >> gpurun_out/r02z_rc.txt
SAN_EPOCHS=3 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py node stream > gpurun_out/r02z_mem_node_stream.log 2>&1; echo "mem node+stream rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r02z_mem_node_stream.log | tail -1)" >> gpurun_out/r02z_rc.txt
cat gpurun_out/r02z_rc.txt; tail -n 4 gpurun_out/r02z_pytest.log | cut -c1-300; tail -n 1 gpurun_out/r02z_smoke.log; tail -c 300 gpurun_out/r02z_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02z_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_python']['value'], d['clocks'], d['gpu_launches'], {k:(v.get('value'), v.get('ms_per_step'), v.get('gpu_launches')) for k,v in d.get('extra_workloads',{}).items()})
d=json.loads(open('gpurun_out/r02z_bench_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'])
"
