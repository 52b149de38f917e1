#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/c2_pytest.log
rm -f gpurun_out/rm_ab.jsonl
python tools/rm_ab.py > gpurun_out/c2_ab.log 2>&1
for v in t256 t64 mb5; do
  RM_AB_CONFIGS=v10,v10+cubes2,v10+cubes3 MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so python tools/rm_ab.py > gpurun_out/c2_ab_$v.log 2>&1
done
cp gpurun_out/rm_ab.jsonl gpurun_out/c2_rm_ab.jsonl
for k in v10+cubes2; do
  RM_ONE=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:raymarch -c 1 -f -o gpurun_out/c2_rm_$k python tools/rm_one.py > gpurun_out/c2_ncu_$k.log 2>&1
  ncu -i gpurun_out/c2_rm_$k.ncu-rep --page source --csv > gpurun_out/c2_rm_${k}_source.csv 2>/dev/null
  ncu -i gpurun_out/c2_rm_$k.ncu-rep --page raw --csv > gpurun_out/c2_rm_${k}_raw.csv 2>/dev/null
done
