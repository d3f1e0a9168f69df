#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learn.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -2 gpurun_out/r2o_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "roofline", "kernel_ms"):
    print(k, d.get(k))
PY
