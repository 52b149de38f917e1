#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/d6_mesh_ab.jsonl
python tools/mesh_ab.py >> gpurun_out/d6_mesh_ab.jsonl 2> gpurun_out/d6_err.log
for v in ilp2 ilp2m5; do
  MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_$v.so python tools/mesh_ab.py >> gpurun_out/d6_mesh_ab.jsonl 2>> gpurun_out/d6_err.log
done
