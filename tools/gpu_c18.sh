#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_group.py -m gpu -q -k "mesh or carve or group or smoke" 2>&1 | tail -6 > gpurun_out/c18_pytest.log
for i in 1 2; do
timeout 300 python tools/kernels_probe.py 2>&1 | grep "K3" >> gpurun_out/c18_k3_bulk.log
MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_nobulk.so timeout 300 python tools/kernels_probe.py 2>&1 | grep "K3" >> gpurun_out/c18_k3_nobulk.log
done
