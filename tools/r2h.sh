#!/bin/bash
mkdir -p gpurun_out
for v in "" "IDQN_NO_ROUNDS=1"; do
  echo "== variant: [$v]"
  for rep in 1 2; do env $v timeout 300 python tools/k_sweep.py --ks 5,1,8 --steps 400 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['K'], d['ms_per_step'], d['kernel_us']['dense_fwd_L3'], d['kernel_us']['dense_dgrad_L3'])"; done
done 2>&1 | tee gpurun_out/r2h_variants.log
