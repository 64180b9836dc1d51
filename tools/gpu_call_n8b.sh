#!/bin/bash
# round-2 GPU call N8b (8 GPUs): the sharded product path on the final build, N = 8 only
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02n8_gpus.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02n8_bench_8gpu.json 2> gpurun_out/r02n8_bench_8gpu.err; echo "bench 8gpu rc=$?" > gpurun_out/r02n8_rc.txt
cat gpurun_out/r02n8_rc.txt; python -c "
import json
d=json.loads(open('gpurun_out/r02n8_bench_8gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['shard_bit_identical'], {k:(round(v['value']), round(v['ms_per_step'],3), round(v['kernel_ms_per_step'],3)) for k,v in d['strong'].items() if isinstance(v,dict)})
"; tail -c 300 gpurun_out/r02n8_bench_8gpu.err
