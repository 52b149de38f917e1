#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/rm_ab.jsonl
for i in 1 2; do
RM_AB_CONFIGS=v10+cubes timeout 300 python tools/rm_ab.py > gpurun_out/c22_ab_default.log 2>&1
RM_AB_CONFIGS=v10+cubes MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_stcs.so timeout 300 python tools/rm_ab.py > gpurun_out/c22_ab_stcs.log 2>&1
done
cp gpurun_out/rm_ab.jsonl gpurun_out/c22_rm_ab.jsonl
