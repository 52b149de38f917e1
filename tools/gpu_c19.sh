#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c19_*
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_group.py -m gpu -q -k "mesh or carve or group or smoke" 2>&1 | tail -6 > gpurun_out/c19_pytest.log
for i in 1 2; do
timeout 300 python tools/kernels_probe.py 2>&1 | grep "K3" >> gpurun_out/c19_k3_bulkstore.log
MESO_SO=$PWD/mesoengine_b200/_variants/libmeso_nostore.so timeout 300 python tools/kernels_probe.py 2>&1 | grep "K3" >> gpurun_out/c19_k3_loopstore.log
done
