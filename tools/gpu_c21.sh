#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/rm_ab.jsonl
RM_AB_CONFIGS=v10+cubes timeout 300 python tools/rm_ab.py > gpurun_out/c21_ab_default.log 2>&1
for v in cs cs256; do
  RM_AB_CONFIGS=v10+cubes MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so timeout 300 python tools/rm_ab.py > gpurun_out/c21_ab_$v.log 2>&1
done
cp gpurun_out/rm_ab.jsonl gpurun_out/c21_rm_ab.jsonl
MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_cs.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "raymarch" 2>&1 | tail -4 > gpurun_out/c21_pytest_cs.log
