#!/bin/bash
mkdir -p gpurun_out
for v in m8 m10 m5; do
  MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so timeout 300 python tools/kernels_probe.py 2>&1 | grep "K3" > gpurun_out/c14_k3_$v.log
done
timeout 300 python tools/kernels_probe.py 2>&1 | grep "K3" > gpurun_out/c14_k3_default.log
