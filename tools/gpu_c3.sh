#!/bin/bash
mkdir -p gpurun_out
MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_dbg.so RM_ONE=v10 RM_ONE_N=256 RM_ONE_W=96 RM_ONE_H=64 timeout 300 python tools/rm_one.py 2>&1 | head -60 > gpurun_out/c3_dbg.log
