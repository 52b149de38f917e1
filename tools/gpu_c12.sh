#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py tests/test_zz_gpu_ref_pin.py -m gpu -q -k "lod or debug or present or block_importance or upload" 2>&1 | tail -15 > gpurun_out/c12_pytest.log
