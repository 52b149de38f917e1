#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/d6_mesh_ab.jsonl
python tools/mesh_ab.py >> gpurun_out/d6_mesh_ab.jsonl 2> gpurun_out/d6_err.log
for v in ax1 ax1m5 ax0m5; do
  MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so python tools/mesh_ab.py >> gpurun_out/d6_mesh_ab.jsonl 2>> gpurun_out/d6_err.log
done
MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_ax1.so timeout 600 python -m pytest tests -m gpu -q -x -k "mesh or carve or remesh" 2>&1 | tail -3 > gpurun_out/d6_pytest_ax1.log
