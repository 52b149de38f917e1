#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/rm_ab.jsonl
timeout 300 python tools/rm_ab.py > gpurun_out/c6_ab_default.log 2>&1
for v in mb5 mb6 t64 t256 pf pfmb5; do
  RM_AB_CONFIGS=v10+cubes MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so timeout 300 python tools/rm_ab.py > gpurun_out/c6_ab_$v.log 2>&1
done
MESO_CUBES_LEVEL=3 RM_AB_CONFIGS=v10+cubes timeout 300 python tools/rm_ab.py > gpurun_out/c6_ab_l3.log 2>&1
MESO_CUBES_LEVEL=1 RM_AB_CONFIGS=v10+cubes timeout 300 python tools/rm_ab.py > gpurun_out/c6_ab_l1.log 2>&1
cp gpurun_out/rm_ab.jsonl gpurun_out/c6_rm_ab.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_cubes.py -m gpu -q 2>&1 | tail -5 > gpurun_out/c6_pytest.log
