#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 0 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_smoke.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_cubes.py tests/test_gpu_group.py -m gpu -q -x -k "not 4096 and not worst_case" > gpurun_out/san_parity.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mesh_quads or occupancy_mips" > gpurun_out/san_race.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_resident.py -m gpu -q -x -k "block_importance or debug or stream_updates" > gpurun_out/san_resident.log 2>&1
