#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c10_pytest.log
timeout 900 python bench.py > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
timeout 300 python tools/kernels_probe.py > gpurun_out/c10_kernels.log 2>&1
