#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/d6_mesh_ab.jsonl
python tools/mesh_ab.py >> gpurun_out/d6_mesh_ab.jsonl 2> gpurun_out/d6_err.log
for v in q2 q3 o8 o7 q2o8; do
  MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so python tools/mesh_ab.py >> gpurun_out/d6_mesh_ab.jsonl 2>> gpurun_out/d6_err.log
done
