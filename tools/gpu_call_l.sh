#!/bin/bash
# round-2 GPU call L: gang kernel v7 (wait-free split rows) A/B against the same build without splitting; stream tests; racecheck of the new shared-memory protocol
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_state.py -q --timeout 300 > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02l_rc.txt
timeout 400 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02l_gang100k_4.log 2>&1; echo "g100k_4 rc=$?" >> gpurun_out/r02l_rc.txt
GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_nosplit/libgnnx.so timeout 400 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02l_gang100k_4_nosplit.log 2>&1; echo "g100k_4 nosplit rc=$?" >> gpurun_out/r02l_rc.txt
timeout 400 python tools/gang_study.py 20000 148 6 0 > gpurun_out/r02l_gang20k_full.log 2>&1; echo "g20k rc=$?" >> gpurun_out/r02l_rc.txt
GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_nosplit/libgnnx.so timeout 400 python tools/gang_study.py 20000 148 6 0 > gpurun_out/r02l_gang20k_full_nosplit.log 2>&1; echo "g20k nosplit rc=$?" >> gpurun_out/r02l_rc.txt
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r02l_bench_c5.json 2> gpurun_out/r02l_bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/r02l_rc.txt
SAN_EPOCHS=3 timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python tools/sanitize_run.py stream > gpurun_out/r02l_race_stream.log 2>&1; echo "race stream rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02l_race_stream.log | tail -1)" >> gpurun_out/r02l_rc.txt
tail -n 5 gpurun_out/r02l_pytest.log | cut -c1-300; for f in r02l_gang100k_4 r02l_gang100k_4_nosplit r02l_gang20k_full r02l_gang20k_full_nosplit; do echo $f; cut -c1-520 gpurun_out/$f.log; done; cat gpurun_out/r02l_rc.txt; python -c "
import json
d=json.loads(open('gpurun_out/r02l_bench_c5.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_at_scale'] and d['parity_at_scale']['rel_l2_max'])
"
