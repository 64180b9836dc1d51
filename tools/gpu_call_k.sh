#!/bin/bash
# round-2 GPU call K: final-build artefacts -- full GPU tests, parity report, contract-only launch list, default bench (driver contract)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02k_rc.txt
timeout 500 python tools/parity_report.py > gpurun_out/r02k_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02k_rc.txt
cp gpurun_out/parity_report.json gpurun_out/r02k_parity_report.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02k_launches_syn1.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/r02k_launchrun.log 2>&1; echo "launch list rc=$?" >> gpurun_out/r02k_rc.txt
python __graft_entry__.py smoke > gpurun_out/r02k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02k_rc.txt
timeout 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?" >> gpurun_out/r02k_rc.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/r02k_rc.txt
cat gpurun_out/r02k_rc.txt; tail -n 4 gpurun_out/r02k_pytest.log | cut -c1-300; tail -n 2 gpurun_out/r02k_smoke.log; tail -c 300 gpurun_out/r02k_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_python']['value'], {k:(v.get('value'), v.get('ms_per_step')) for k,v in d.get('extra_workloads',{}).items()})
d=json.loads(open('gpurun_out/r02k_bench_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'], d.get('cpu_baseline',{}).get('cores'))
"
