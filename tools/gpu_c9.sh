#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_cubes.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q 2>&1 | tail -12 > gpurun_out/c9_pytest.log
timeout 300 python tools/kernels_probe.py > gpurun_out/c9_kernels.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c9_launches.csv python tools/mesh_probe.py > /dev/null 2>&1
