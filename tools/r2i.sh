#!/bin/bash
# round-2 final 1-GPU pass: tests, bench line, timelines, acting latency, training-loop overlap
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2i_smoke.log
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -2 gpurun_out/r2i_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "step_hbm_frac", "clocks", "e2e", "strong_k8", "other_configs_us_per_step", "prioritized_replay", "cpu_baseline", "roofline", "kernel_ms"):
    print(k, d.get(k))
PY
for k in 1 5 8; do timeout 300 python tools/kernel_timeline.py $k > gpurun_out/r2i_timeline_k$k.txt 2>&1; tail -1 gpurun_out/r2i_timeline_k$k.txt; done
timeout 300 python tools/best_action_latency.py 2>&1 | tee gpurun_out/r2i_best_action.log
timeout 600 python tools/train_overlap.py 5 0 150 300 2>&1 | tee gpurun_out/r2i_train_overlap.log
timeout 300 python tools/replay_add_latency.py 2>&1 | tee gpurun_out/r2i_replay_add.log
