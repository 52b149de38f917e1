#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/c11_pytest.log
timeout 900 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
