#!/bin/bash
# two-level greedy merge: parity of every test that meshes, then per-kernel times of the mesh passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or carve or group or remesh or stream or window or pool or lod or host_sample" 2>&1 | tail -12 > gpurun_out/d1_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/d1_launches_kernels.csv python tools/kernels_probe.py > gpurun_out/d1_kernels_probe.log 2>&1
