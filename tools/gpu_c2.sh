#!/bin/bash
# Round 2, GPU call 2: locate the illegal access of the v10 walk (compute-sanitizer, -lineinfo), small then full size.
mkdir -p gpurun_out
RM_ONE=v10 RM_ONE_N=1024 RM_ONE_W=1920 RM_ONE_H=1080 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/rm_one.py > gpurun_out/c2_san_1024.log 2>&1
RM_ONE=v10 RM_ONE_N=4096 RM_ONE_W=3840 RM_ONE_H=2160 timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/rm_one.py > gpurun_out/c2_san_4096.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_cubes.py -m gpu -q 2>&1 | tail -40 > gpurun_out/c2_pytest_v10.log
