#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/partition_probe.py > gpurun_out/c17_partition.json 2> gpurun_out/c17_partition.err
