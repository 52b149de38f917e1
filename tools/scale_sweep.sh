#!/bin/bash
# 1/2/4/8-GPU sweep of bench.py on one box (what the driver does at round end); results in gpurun_out/scale_*.json
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 80 --warmup 8 --no-mesh --no-cpu 2>gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 80 --warmup 8 2>gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
  fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_$N.json"))
    print("N=$N value=%.0f Mrays/s ms=%.3f kernel_ms=%.3f e2e=%.0f gather=%s verified=%s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["config"].get("gather"), d["config"].get("gather_verified_equal_to_1gpu_frame")))
except Exception as e:
    print("N=$N failed:", e); print(open("gpurun_out/scale_$N.err").read()[-1500:])
PY
done
