#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/c4_pytest.log
rm -f gpurun_out/rm_ab.jsonl
timeout 600 python tools/rm_ab.py > gpurun_out/c4_ab.log 2>&1
cp gpurun_out/rm_ab.jsonl gpurun_out/c4_rm_ab.jsonl
