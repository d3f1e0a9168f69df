#!/bin/bash
# A/B of environment knobs: tools/sweep_env.sh <tag> "<ENV=.. ENV=..> <flags>" ...
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  echo "== $spec" | tee -a gpurun_out/${tag}_sweep.log
  f=${spec##* }; envs=${spec% *}; [ "$envs" == "$spec" ] && envs=""
  env $envs timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --flags $f 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {k:round(v*1e3,1) for k,v in d['kernel_ms'].items()})" | tee -a gpurun_out/${tag}_sweep.log
done
