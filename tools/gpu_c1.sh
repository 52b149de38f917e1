#!/bin/bash
# Round 2, GPU call 1: all GPU tests on the default (v10) walk and on v8, A/B of the walks and of the launch shapes,
# the default bench line, one ncu --set full capture of the v10 walk with per-line source counters.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/c1_pytest.log
rm -f gpurun_out/rm_ab.jsonl
timeout 600 python tools/rm_ab.py > gpurun_out/c1_ab_default.log 2>&1
for v in mb5 mb6 t256 t64; do
  RM_AB_CONFIGS=v10,v10+cubes1,v10+cubes3 MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so timeout 300 python tools/rm_ab.py > gpurun_out/c1_ab_$v.log 2>&1
done
cp gpurun_out/rm_ab.jsonl gpurun_out/c1_rm_ab.jsonl
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
for k in v10 v10+cubes1; do
  RM_ONE=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:raymarch -c 1 -f -o gpurun_out/c1_rm_$k python tools/rm_one.py > gpurun_out/c1_ncu_$k.log 2>&1
  ncu -i gpurun_out/c1_rm_$k.ncu-rep --page source --csv > gpurun_out/c1_rm_${k}_source.csv 2>/dev/null
  ncu -i gpurun_out/c1_rm_$k.ncu-rep --page raw --csv > gpurun_out/c1_rm_${k}_raw.csv 2>/dev/null
done
MESO_RM_KERNEL=v8 timeout 600 python -m pytest tests/test_zz_gpu_cubes.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 > gpurun_out/c1_pytest_v8.log
ls -la gpurun_out > gpurun_out/c1_ls.txt
