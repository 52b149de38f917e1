#!/bin/bash
mkdir -p gpurun_out
MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_dbg2.so RM_ONE=v10 RM_ONE_N=256 RM_ONE_W=96 RM_ONE_H=64 timeout 300 python tools/rm_one.py 2>&1 | grep -v "bad ci" | head -80 > gpurun_out/c3_dbg2.log
MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_noslow.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "raymarch_sphere_voxel_256 or raymarch_terrain" 2>&1 | tail -5 > gpurun_out/c3_noslow.log
