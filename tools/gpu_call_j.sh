#!/bin/bash
# round-2 GPU call J: gang kernel v6 (mid rows in rounds), model forward on device -- tests, study, c5 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02j_rc.txt
timeout 400 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02j_gang100k_4.log 2>&1; echo "g100k_4 rc=$?" >> gpurun_out/r02j_rc.txt
timeout 400 python tools/gang_study.py 20000 148 6 0,37 > gpurun_out/r02j_gang20k_full.log 2>&1; echo "g20k_full rc=$?" >> gpurun_out/r02j_rc.txt
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r02j_bench_c5.json 2> gpurun_out/r02j_bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/r02j_rc.txt
tail -n 12 gpurun_out/r02j_pytest.log | cut -c1-300; for f in r02j_gang100k_4 r02j_gang20k_full; do echo $f; cut -c1-520 gpurun_out/$f.log; done; cat gpurun_out/r02j_rc.txt; python -c "
import json
d=json.loads(open('gpurun_out/r02j_bench_c5.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_at_scale'] and d['parity_at_scale']['rel_l2_max'])
"
