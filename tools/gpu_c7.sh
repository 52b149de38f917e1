#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/rm_ab.jsonl
RM_AB_CONFIGS=v10+cubes timeout 300 python tools/rm_ab.py > gpurun_out/c7_ab_default.log 2>&1
for v in mb5 t64 t64mb5 t32 pf pft64; do
  RM_AB_CONFIGS=v10+cubes MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so timeout 300 python tools/rm_ab.py > gpurun_out/c7_ab_$v.log 2>&1
done
cp gpurun_out/rm_ab.jsonl gpurun_out/c7_rm_ab.jsonl
