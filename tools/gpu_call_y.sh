#!/bin/bash
# round-2 GPU call Y: build with the automatic cluster class, max carveout, 512-thread extraction, lazy plan fetch -- full GPU tests, smoke,
# default bench, cluster study, plan timing, racecheck + memcheck of the cluster class
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02y_rc.txt
python __graft_entry__.py smoke > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02y_rc.txt
timeout 900 python bench.py > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?" >> gpurun_out/r02y_rc.txt
GNNX_STUDY_TAG=_y timeout 600 python tools/cluster_study.py syn1 1,8,24,88,130,175,260,350,0 1,0 > gpurun_out/r02y_cluster_study.log 2>&1; echo "study rc=$?" >> gpurun_out/r02y_rc.txt
timeout 200 python tools/plan_time.py > gpurun_out/r02y_plan_time.json 2> gpurun_out/r02y_plan_time.err; echo "plan rc=$?" >> gpurun_out/r02y_rc.txt
SAN_EPOCHS=3 timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py cluster > gpurun_out/r02y_race_cluster.log 2>&1; echo "race cluster rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02y_race_cluster.log | tail -1)" >> gpurun_out/r02y_rc.txt
SAN_EPOCHS=3 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py cluster node misc > gpurun_out/r02y_mem.log 2>&1; echo "mem rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r02y_mem.log | tail -1)" >> gpurun_out/r02y_rc.txt
cat gpurun_out/r02y_rc.txt; tail -n 4 gpurun_out/r02y_pytest.log | cut -c1-300; tail -n 1 gpurun_out/r02y_smoke.log; tail -c 300 gpurun_out/r02y_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02y_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['e2e_python']['value'], d['clocks'], d['gpu_launches'], {k:(v.get('value'), v.get('ms_per_step'), v.get('gpu_launches')) for k,v in d.get('extra_workloads',{}).items()})
for r in json.load(open('gpurun_out/cluster_study_syn1_y.json')): print(r['batch'], r['cluster'], round(r['kernel_ms_min'],3), r['class_counts'], r['plan_cluster'], 'end', r['class_end_ms'])
print(open('gpurun_out/r02y_plan_time.json').read())
"
