#!/bin/bash
# round-2 GPU pass E (1 GPU): tests, acting latency, K sweep, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2e_pytest.log
timeout 300 python tools/best_action_latency.py 2>&1 | tee gpurun_out/r2e_best_action.log
timeout 600 python tools/k_sweep.py --ks 1,5,8 --steps 300 > gpurun_out/r2e_ksweep.jsonl 2> gpurun_out/r2e_ksweep.err; cut -c1-110 gpurun_out/r2e_ksweep.jsonl
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -2 gpurun_out/r2e_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "step_hbm_frac", "clocks", "e2e", "strong_k8", "other_configs_us_per_step", "prioritized_replay", "cpu_baseline", "roofline"):
    print(k, d.get(k))
PY
