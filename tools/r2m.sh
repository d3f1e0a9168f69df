#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for K in 5 1 8; do timeout 200 python tools/kernel_timeline.py $K 2>&1 | tail -18 > gpurun_out/r2m_timeline_k$K.txt; tail -1 gpurun_out/r2m_timeline_k$K.txt; done
timeout 300 python tools/k_sweep.py --ks 1,2,3,5,8 2>&1 | tail -5 | tee gpurun_out/r2m_ksweep.txt | cut -c1-70
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -2 gpurun_out/r2m_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], "roofline", d["roofline"])
print({k: d[k] for k in d if k not in ("e2e", "roofline", "config")})
PY
