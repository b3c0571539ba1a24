#!/bin/bash
# 8-GPU run: BASELINE configs[4] at full size, then the 8-GPU bench line
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r2_8gpu_host.txt; nproc >> gpurun_out/r2_8gpu_host.txt; nvidia-smi topo -m >> gpurun_out/r2_8gpu_host.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tests/full_configs.py 5 > gpurun_out/r2_cfg5_8gpu.jsonl 2> gpurun_out/r2_cfg5_8gpu.err; echo "rc=$?" >> gpurun_out/r2_cfg5_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "rc=$?" >> gpurun_out/r2_bench_8gpu.err
tail -5 gpurun_out/r2_cfg5_8gpu.err; cat gpurun_out/r2_cfg5_8gpu.jsonl; tail -3 gpurun_out/r2_bench_8gpu.err; grep "^{" gpurun_out/r2_bench_8gpu.json | cut -c1-400
