#!/bin/bash
# Round 2, GPU call 5: all GPU tests with the forward cubes as the default walk; bench line; ncu --set full of the default walk.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/c5_pytest.log
timeout 900 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
for k in v10+cubes; do
  RM_ONE=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:raymarch -c 1 -f -o gpurun_out/c5_rm python tools/rm_one.py > gpurun_out/c5_ncu.log 2>&1
  ncu -i gpurun_out/c5_rm.ncu-rep --page source --csv > gpurun_out/c5_rm_source.csv 2>/dev/null
  ncu -i gpurun_out/c5_rm.ncu-rep --page raw --csv > gpurun_out/c5_rm_raw.csv 2>/dev/null
done
