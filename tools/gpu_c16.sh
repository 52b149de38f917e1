#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_zz_gpu_ref_pin.py tests/test_gpu_resident.py -m gpu -q -k "occupancy or instances or reference or window or smoke or upload or terrain" 2>&1 | tail -8 > gpurun_out/c16_pytest.log
timeout 300 python tools/kernels_probe.py > gpurun_out/c16_kernels.log 2>&1
