#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for K in 5 1; do
  echo "== K=$K wgrad0 on main"; timeout 200 python tools/kernel_timeline.py $K 2>&1 | tail -24 | tee gpurun_out/r2k_timeline_k$K.txt | grep -i "step\|total\|wgrad\|adam"
done
timeout 300 python tools/k_sweep.py --ks 1,2,3,5,8 2>&1 | tail -12 | cut -c1-70 | tee gpurun_out/r2k_ksweep.txt
