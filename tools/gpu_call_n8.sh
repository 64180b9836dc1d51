#!/bin/bash
# round-2 GPU call N8 (8 GPUs): the sharded product path at the full node scale -- bench lines at N = 8 and N = 4
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02n_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02n_bench_8gpu.json 2> gpurun_out/r02n_bench_8gpu.err; echo "bench 8gpu rc=$?" > gpurun_out/r02n_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02n_bench_4gpu.json 2> gpurun_out/r02n_bench_4gpu.err; echo "bench 4gpu rc=$?" >> gpurun_out/r02n_rc.txt
cat gpurun_out/r02n_rc.txt; python -c "
import json
for f in ('gpurun_out/r02n_bench_8gpu.json','gpurun_out/r02n_bench_4gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['shard_bit_identical'], {k:(v['value'], v['ms_per_step']) for k,v in d['strong'].items() if isinstance(v,dict)})
    except Exception as e: print(f, 'ERR', e)
"; tail -c 400 gpurun_out/r02n_bench_8gpu.err
