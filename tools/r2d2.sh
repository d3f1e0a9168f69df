#!/bin/bash
# round-2 final N-GPU pass: sharded check (peer-memory exchange) and bench --gpus N
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -8 | tee gpurun_out/r2d_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 300 --warmup 20 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err
tail -3 gpurun_out/r2d_bench_n$N.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "dropin", d["e2e"]["blocking_call_value"])
    print("strong_k8", d["strong_k8"]); print("parity", d.get("sharded_parity")); print("roofline", d["roofline"])
except Exception as e:
    print("no bench line:", e)
PY
