#!/bin/bash
# Round 2, GPU call 8: all GPU tests (group API, slabs, callback sample), profile of the final raymarch kernel, of the mesh
# and occupancy kernels, kernel timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/c8_pytest.log
RM_ONE=v10+cubes timeout 600 ncu --set full --clock-control none --import-source on -k regex:raymarch -c 1 -f -o gpurun_out/c8_rm python tools/rm_one.py > gpurun_out/c8_ncu_rm.log 2>&1
ncu -i gpurun_out/c8_rm.ncu-rep --page source --csv > gpurun_out/c8_rm_source.csv 2>/dev/null
ncu -i gpurun_out/c8_rm.ncu-rep --page raw --csv > gpurun_out/c8_rm_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_bricks|mesh_worklist" -c 2 -f -o gpurun_out/c8_mesh python tools/mesh_probe.py > gpurun_out/c8_ncu_mesh.log 2>&1
ncu -i gpurun_out/c8_mesh.ncu-rep --page source --csv > gpurun_out/c8_mesh_source.csv 2>/dev/null
ncu -i gpurun_out/c8_mesh.ncu-rep --page raw --csv > gpurun_out/c8_mesh_raw.csv 2>/dev/null
timeout 300 python tools/kernels_probe.py > gpurun_out/c8_kernels.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c8_launches.csv python tools/kernels_probe.py > /dev/null 2>&1
