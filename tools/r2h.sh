#!/bin/bash
mkdir -p gpurun_out
for v in "" "IDQN_NO_PRIORITY=1" "IDQN_NO_RESIDENT=1" "IDQN_NO_PRIORITY=1 IDQN_NO_RESIDENT=1" "IDQN_NO_PAIR=1"; do
  echo "== variant: [$v]"
  env $v timeout 300 python tools/k_sweep.py --ks 5 --steps 400 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_us'])"
  env $v timeout 300 python tools/k_sweep.py --ks 5 --steps 400 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
done 2>&1 | tee gpurun_out/r2h_variants.log
