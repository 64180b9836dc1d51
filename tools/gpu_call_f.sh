#!/bin/bash
# round-2 GPU call F: gang kernel v4 (level-order feature copy: one bulk copy per tile; unrolled dL/dsF reduction) -- tests, study, c5 bench, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r02f_rc.txt
timeout 400 python tools/gang_study.py 100000 4 6 0,37 > gpurun_out/r02f_gang100k_4.log 2>&1; echo "g100k_4 rc=$?" >> gpurun_out/r02f_rc.txt
timeout 600 python tools/gang_study.py 100000 148 4 74,148 > gpurun_out/r02f_gang100k_full.log 2>&1; echo "g100k_full rc=$?" >> gpurun_out/r02f_rc.txt
timeout 400 python tools/gang_study.py 20000 148 6 0,37 > gpurun_out/r02f_gang20k_full.log 2>&1; echo "g20k_full rc=$?" >> gpurun_out/r02f_rc.txt
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/r02f_bench_c5.json 2> gpurun_out/r02f_bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/r02f_rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:explain_gang_kernel -c 1 -f -o gpurun_out/r02f_gang python tools/ncu_target.py c5 100000 1 3 > gpurun_out/r02f_ncu_gang.log 2>&1; echo "ncu gang rc=$?" >> gpurun_out/r02f_rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:explain_node_kernel -c 6 -f -o gpurun_out/r02f_node python tools/ncu_target.py syn1 > gpurun_out/r02f_ncu_node.log 2>&1; echo "ncu node rc=$?" >> gpurun_out/r02f_rc.txt
tail -n 6 gpurun_out/r02f_pytest.log | cut -c1-300; for f in r02f_gang100k_4 r02f_gang100k_full r02f_gang20k_full; do echo $f; cut -c1-520 gpurun_out/$f.log; done; cat gpurun_out/r02f_rc.txt; tail -c 300 gpurun_out/r02f_bench_c5.err; tail -n 3 gpurun_out/r02f_ncu_gang.log gpurun_out/r02f_ncu_node.log; ls -la gpurun_out/*.ncu-rep
