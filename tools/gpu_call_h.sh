#!/bin/bash
# round-2 GPU call H (2 GPUs): the sharded product path over NCCL -- test + bench lines
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -q --timeout 300 > gpurun_out/r02h_pytest_dist.log 2>&1; echo "pytest dist rc=$?" > gpurun_out/r02h_rc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err; echo "bench 2gpu rc=$?" >> gpurun_out/r02h_rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02h_bench_2gpu_ref.json 2> gpurun_out/r02h_bench_2gpu_ref.err; echo "bench 2gpu ref rc=$?" >> gpurun_out/r02h_rc.txt
cat gpurun_out/r02h_rc.txt; tail -n 5 gpurun_out/r02h_pytest_dist.log | cut -c1-300; tail -c 1800 gpurun_out/r02h_bench_2gpu.json; tail -c 500 gpurun_out/r02h_bench_2gpu.err; tail -c 400 gpurun_out/r02h_bench_2gpu_ref.json
