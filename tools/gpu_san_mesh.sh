#!/bin/bash
# compute-sanitizer over the mesh / edit paths (memcheck, then racecheck on the shared-memory staging of passes B and C)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -q -x -k "mesh or carve or remesh or group" > gpurun_out/san_mesh_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mesh_quads or mesh_brick_level or mesh_worst" > gpurun_out/san_mesh_racecheck.log 2>&1
