#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_zz_gpu_ref_pin.py tests/test_gpu_host_sample.py -m gpu -q 2>&1 | tail -6 > gpurun_out/c20_pytest.log
timeout 300 python tools/resident_probe.py > gpurun_out/c20_resident.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c20_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c20_smoke.log 2>&1
