#!/bin/bash
# round-2 GPU call V: FINAL build (sub-warp k-hop rows, all-warp S gather in the gang kernel) -- gang A/B, full GPU tests, smoke, default bench, contract launch list
mkdir -p gpurun_out; rm -f gpurun_out/r02v_rc.txt
for n in default sub default sub; do
  unset GNNX_LIB_PATH
  if [ $n != default ]; then export GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_$n/libgnnx.so; fi
  timeout 300 python tools/gang_study.py 100000 4 6 0 > gpurun_out/r02v_gang_$n.log 2>&1; echo "gang $n rc=$?" >> gpurun_out/r02v_rc.txt
  echo "== $n"; grep "^{" gpurun_out/r02v_gang_$n.log | tail -1 | cut -c1-420
done
unset GNNX_LIB_PATH
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_rc.txt
python __graft_entry__.py smoke > gpurun_out/r02v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02v_rc.txt
timeout 900 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?" >> gpurun_out/r02v_rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02v_launches_syn1.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/r02v_launchrun.log 2>&1; echo "launch list rc=$?" >> gpurun_out/r02v_rc.txt
cat gpurun_out/r02v_rc.txt; tail -n 3 gpurun_out/r02v_pytest.log | cut -c1-300; tail -n 1 gpurun_out/r02v_smoke.log; tail -c 300 gpurun_out/r02v_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['e2e_python']['value'], d['e2e_python'].get('single_explain_call_ms'), d['e2e_python'].get('single_explain_call_ms_latency_mode'), d['clocks'], d['gpu_launches'], {k:(v.get('value'), v.get('ms_per_step'), v.get('roofline',{}).get('frac')) for k,v in d.get('extra_workloads',{}).items()})
"
