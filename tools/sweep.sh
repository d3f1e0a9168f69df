#!/bin/bash
# quick A/B of engine flags on the GPU box: tools/sweep.sh <tag> <flags...>
tag=$1; shift
mkdir -p gpurun_out
for f in "$@"; do
  echo "== flags $f" | tee -a gpurun_out/${tag}_sweep.log
  timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --flags $f 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {k:round(v*1e3,1) for k,v in d['kernel_ms'].items()})" | tee -a gpurun_out/${tag}_sweep.log
done
