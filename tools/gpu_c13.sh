#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_group.py -m gpu -q -k "mesh or carve or group or smoke" 2>&1 | tail -15 > gpurun_out/c13_pytest.log
timeout 300 python tools/kernels_probe.py > gpurun_out/c13_kernels.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_bricks" -c 1 -f -o gpurun_out/c13_mesh python tools/mesh_probe.py > gpurun_out/c13_ncu_mesh.log 2>&1
ncu -i gpurun_out/c13_mesh.ncu-rep --page source --csv > gpurun_out/c13_mesh_source.csv 2>/dev/null
ncu -i gpurun_out/c13_mesh.ncu-rep --page raw --csv > gpurun_out/c13_mesh_raw.csv 2>/dev/null
