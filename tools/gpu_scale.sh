#!/bin/bash
# scaling pass with the final kernels: N = $1
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/scale_N$N.json 2> gpurun_out/scale_N$N.err
