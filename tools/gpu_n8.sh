#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 40 --warmup 8 --no-group > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 40 --warmup 8 --frames-in-flight 8 --no-group --no-mesh > gpurun_out/n8_bench_r8.json 2> gpurun_out/n8_bench_r8.err
