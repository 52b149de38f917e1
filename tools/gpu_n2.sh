#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q 2>&1 | tail -4 > gpurun_out/n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
