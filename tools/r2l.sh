#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 0.3" "48 0.4" "48 0.5" "64 0.5" "64 0.6" "64 0.7" "80 0.7" "80 0.8" "96 0.9"; do
  set -- $cfg
  echo "== IDQN_WG_OVERLAP=$1 IDQN_WG_FRAC=$2"; IDQN_WG_OVERLAP=$1 IDQN_WG_FRAC=$2 timeout 300 python tools/k_sweep.py --ks 1,3,5,8 2>&1 | tail -4 | cut -c1-70
done
