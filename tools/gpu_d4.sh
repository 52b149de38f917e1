#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or carve or group or remesh" 2>&1 | tail -12 > gpurun_out/d4_pytest.log
timeout 600 python tools/kernels_probe.py > gpurun_out/d4_kernels_probe.log 2>&1
timeout 900 python bench.py --no-cpu > gpurun_out/d4_bench.json 2> gpurun_out/d4_bench.err
