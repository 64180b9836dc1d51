#!/bin/bash
# round-2 GPU call T: scheduling-knob sweep of the syn1 batch; bench line with the child-process clock sampler
mkdir -p gpurun_out
timeout 600 python bench.py --no-extra > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?" > gpurun_out/r02t_rc.txt
timeout 900 python tools/tune_syn1.py > gpurun_out/r02t_tune.log 2>&1; echo "tune rc=$?" >> gpurun_out/r02t_rc.txt
cat gpurun_out/r02t_rc.txt; python -c "
import json
d=json.loads(open('gpurun_out/r02t_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])
"; sort -t: -k2 gpurun_out/r02t_tune.log | head -50 | cut -c1-200
