#!/bin/bash
# 1-GPU validation pass: all GPU tests, the default bench line, the reference arm, launch lists of the probe and the smoke run, ncu of the mesh kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/val_pytest.log
timeout 900 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/val_bench_reference.json 2> gpurun_out/val_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/val_launches_kernels.csv python tools/kernels_probe.py > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/val_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mesh_bricks" -c 1 -f -o gpurun_out/val_mesh python tools/mesh_probe.py > gpurun_out/val_ncu_mesh.log 2>&1
ncu -i gpurun_out/val_mesh.ncu-rep --page source --csv > gpurun_out/val_mesh_source.csv 2>/dev/null
ncu -i gpurun_out/val_mesh.ncu-rep --page raw --csv > gpurun_out/val_mesh_raw.csv 2>/dev/null
