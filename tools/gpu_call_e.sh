#!/bin/bash
# round-2 GPU call E: gang kernel v3 (row-parallel short rows, L2 policies, redundant B2, unroll 8) -- tests + A/B study + bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02e_rc.txt
timeout 400 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02e_gang100k_4.log 2>&1; echo "g100k_4 rc=$?" >> gpurun_out/r02e_rc.txt
GNNX_LIB_PATH=$PWD/gnn-model-explainer_b200/gnnx/lib_un4/libgnnx.so timeout 400 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02e_gang100k_4_un4.log 2>&1; echo "g100k_4 un4 rc=$?" >> gpurun_out/r02e_rc.txt
timeout 600 python tools/gang_study.py 100000 148 4 -1,37,74,148 > gpurun_out/r02e_gang100k_full.log 2>&1; echo "g100k_full rc=$?" >> gpurun_out/r02e_rc.txt
timeout 400 python tools/gang_study.py 20000 148 6 -1,0,18,37 > gpurun_out/r02e_gang20k_full.log 2>&1; echo "g20k_full rc=$?" >> gpurun_out/r02e_rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?" >> gpurun_out/r02e_rc.txt
tail -n 8 gpurun_out/r02e_pytest.log | cut -c1-300; for f in r02e_gang100k_4 r02e_gang100k_4_un4 r02e_gang100k_full r02e_gang20k_full; do echo $f; cut -c1-520 gpurun_out/$f.log; done; cat gpurun_out/r02e_rc.txt; tail -c 400 gpurun_out/r02e_bench.err
