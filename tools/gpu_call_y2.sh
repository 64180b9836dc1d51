#!/bin/bash
# round-2 GPU call Y2: strict default (no clusters), plan sync restored, 512-thread extraction, max carveout -- full GPU tests, smoke, default bench, plan timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02y2_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02y2_rc.txt
python __graft_entry__.py smoke > gpurun_out/r02y2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02y2_rc.txt
timeout 900 python bench.py > gpurun_out/r02y2_bench.json 2> gpurun_out/r02y2_bench.err; echo "bench rc=$?" >> gpurun_out/r02y2_rc.txt
timeout 200 python tools/plan_time.py > gpurun_out/r02y2_plan_time.json 2> gpurun_out/r02y2_plan_time.err; echo "plan rc=$?" >> gpurun_out/r02y2_rc.txt
GNNX_HOST_TIMING=1 timeout 200 python tools/plan_time.py 2>&1 | grep "gnnx\]" | tail -2
cat gpurun_out/r02y2_rc.txt; tail -n 4 gpurun_out/r02y2_pytest.log | cut -c1-300; tail -n 1 gpurun_out/r02y2_smoke.log; tail -c 300 gpurun_out/r02y2_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02y2_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['e2e_python']['value'], d['e2e_python'].get('single_explain_call_ms'), d['e2e_python'].get('single_explain_call_ms_latency_mode'), d['clocks'], d['gpu_launches'], {k:(v.get('value'), v.get('ms_per_step'), v.get('gpu_launches')) for k,v in d.get('extra_workloads',{}).items()})
print(open('gpurun_out/r02y2_plan_time.json').read())
"
