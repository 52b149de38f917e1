#!/bin/bash
# A/B of K3 builds on one box: the shipped library, then every mesoengine_b200/_variants/libmeso_*.so (tools/build_variant.sh);
# equal fingerprints = equal quad lists.   -> gpurun_out/mesh_ab.jsonl
mkdir -p gpurun_out
: > gpurun_out/mesh_ab.jsonl
python tools/mesh_ab.py >> gpurun_out/mesh_ab.jsonl 2> gpurun_out/mesh_ab.err
for so in mesoengine_b200/_variants/libmeso_*.so; do
  [ -e "$so" ] || continue
  MESO_SO=$PWD/$so python tools/mesh_ab.py >> gpurun_out/mesh_ab.jsonl 2>> gpurun_out/mesh_ab.err
done
