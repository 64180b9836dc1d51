#!/bin/bash
# round-2 GPU call I: steady-state ncu of the gang kernel (1 task, 16 epochs), launch list of the bench command, headline bench
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:explain_gang_kernel -c 1 -f -o gpurun_out/r02i_gang16 python tools/ncu_target.py c5 100000 1 16 > gpurun_out/r02i_ncu_gang.log 2>&1; echo "ncu gang rc=$?" > gpurun_out/r02i_rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02i_launchrun.log 2>&1; echo "launch list rc=$?" >> gpurun_out/r02i_rc.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -k "batch_of_graphs or dropin" > gpurun_out/r02i_pytest_sel.log 2>&1; echo "pytest sel rc=$?" >> gpurun_out/r02i_rc.txt
timeout 900 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?" >> gpurun_out/r02i_rc.txt
cat gpurun_out/r02i_rc.txt; tail -n 3 gpurun_out/r02i_ncu_gang.log; wc -l gpurun_out/r02i_launches.csv; tail -c 600 gpurun_out/r02i_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_python']['value'] if 'e2e_python' in d else None, {k:(v.get('value'), v.get('ms_per_step')) for k,v in d.get('extra_workloads',{}).items()})
"
