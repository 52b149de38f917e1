#!/bin/bash
# GPU call 1 of round 2: all GPU tests (cubes un-skipped), A/B of the raymarch walks (v8 / v10, cube levels, 48 / 64 registers),
# one ncu --set full capture per walk with per-line source counters, and the launch list of the smoke run.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/c1_pytest.log
MESO_RM_KERNEL=v8 python -m pytest tests/test_zz_gpu_cubes.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -25 > gpurun_out/c1_pytest_v8.log
rm -f gpurun_out/rm_ab.jsonl
python tools/rm_ab.py > gpurun_out/c1_ab_mb5.log 2>&1
RM_AB_CONFIGS=v10,v10+cubes1,v10+cubes3 MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_mb4.so python tools/rm_ab.py > gpurun_out/c1_ab_mb4.log 2>&1
for k in v8 v10 v10+cubes3; do
  RM_ONE=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:raymarch -c 1 -f -o gpurun_out/c1_rm_$k python tools/rm_one.py > gpurun_out/c1_ncu_$k.log 2>&1
  ncu -i gpurun_out/c1_rm_$k.ncu-rep --page source --csv > gpurun_out/c1_rm_${k}_source.csv 2>/dev/null
  ncu -i gpurun_out/c1_rm_$k.ncu-rep --page raw --csv > gpurun_out/c1_rm_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out > gpurun_out/c1_ls.txt
