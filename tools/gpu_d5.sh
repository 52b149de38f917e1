#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_bricks" -c 1 -f -o gpurun_out/d5_mesh python tools/mesh_probe.py > gpurun_out/d5_ncu_mesh.log 2>&1
ncu -i gpurun_out/d5_mesh.ncu-rep --page source --csv > gpurun_out/d5_mesh_source.csv 2>/dev/null
ncu -i gpurun_out/d5_mesh.ncu-rep --page raw --csv > gpurun_out/d5_mesh_raw.csv 2>/dev/null
