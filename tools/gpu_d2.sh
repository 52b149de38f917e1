#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or carve or group or remesh" 2>&1 | tail -12 > gpurun_out/d2_pytest.log
timeout 600 python tools/kernels_probe.py > gpurun_out/d2_kernels_probe.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/d2_launches_kernels.csv python tools/kernels_probe.py > /dev/null 2>&1
