#!/bin/bash
# round-2 GPU call G: gang kernel v5 (interleaved MMA chains, recomputed sigmoid) -- tests, study, benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02g_rc.txt
timeout 400 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02g_gang100k_4.log 2>&1; echo "g100k_4 rc=$?" >> gpurun_out/r02g_rc.txt
timeout 400 python tools/gang_study.py 20000 148 6 0,37 > gpurun_out/r02g_gang20k_full.log 2>&1; echo "g20k_full rc=$?" >> gpurun_out/r02g_rc.txt
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r02g_bench_c5.json 2> gpurun_out/r02g_bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/r02g_rc.txt
timeout 600 python bench.py --workload c5 --c5-n 20000 --steps 2 --warmup 1 > gpurun_out/r02g_bench_c5_n20k.json 2> gpurun_out/r02g_bench_c5_n20k.err; echo "bench c5 20k rc=$?" >> gpurun_out/r02g_rc.txt
tail -n 6 gpurun_out/r02g_pytest.log | cut -c1-300; for f in r02g_gang100k_4 r02g_gang20k_full; do echo $f; cut -c1-520 gpurun_out/$f.log; done; cat gpurun_out/r02g_rc.txt; tail -c 300 gpurun_out/r02g_bench_c5.err; python -c "
import json
for f in ('gpurun_out/r02g_bench_c5.json','gpurun_out/r02g_bench_c5_n20k.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_at_scale'] and d['parity_at_scale']['rel_l2_max'])
"
