#!/bin/bash
mkdir -p gpurun_out
for v in "" "IDQN_WGRAD_LATE=1"; do
  echo "== variant: [$v]"
  env $v timeout 300 python tools/k_sweep.py --ks 5,1 --steps 400 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['K'], d['ms_per_step'])"
  env $v timeout 300 python tools/kernel_timeline.py 5 2>&1 | tail -9
done 2>&1 | tee gpurun_out/r2h_variants.log
