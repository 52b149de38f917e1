#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 60 --warmup 8 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 60 --warmup 8 --rendezvous nccl --no-group --no-mesh > gpurun_out/n8_bench_nccl.json 2> gpurun_out/n8_bench_nccl.err
