#!/bin/bash
# round-2 GPU call N4 (4 GPUs): the sharded product path on the final build (memoised shard layout, latency-mode strong line)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02n4_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02n4_bench_4gpu.json 2> gpurun_out/r02n4_bench_4gpu.err; echo "bench 4gpu rc=$?" > gpurun_out/r02n4_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02n4_bench_2gpu.json 2> gpurun_out/r02n4_bench_2gpu.err; echo "bench 2gpu rc=$?" >> gpurun_out/r02n4_rc.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r02n4_dist_tests.log 2>&1; echo "dist tests rc=$?" >> gpurun_out/r02n4_rc.txt; tail -2 gpurun_out/r02n4_dist_tests.log
cat gpurun_out/r02n4_rc.txt; python -c "
import json
for f in ('gpurun_out/r02n4_bench_4gpu.json','gpurun_out/r02n4_bench_2gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['shard_bit_identical'], {k:(round(v['value']), round(v['ms_per_step'],3), round(v['kernel_ms_per_step'],3)) for k,v in d['strong'].items() if isinstance(v,dict)})
    except Exception as e: print(f, 'ERR', e)
"; tail -c 300 gpurun_out/r02n4_bench_4gpu.err
