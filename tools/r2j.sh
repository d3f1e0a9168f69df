#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learn.py -x -q -m gpu 2>&1 | tail -3
for K in 5 3 2 1; do
  echo "== K=$K unit deps"; timeout 200 python tools/kernel_timeline.py $K 2>&1 | tail -24 | tee gpurun_out/r2j_timeline_k$K.txt | grep -i "step\|total\|img_"
  echo "== K=$K whole-grid waits"; IDQN_NO_UNIT_DEPS=1 timeout 200 python tools/kernel_timeline.py $K 2>&1 | tail -24 | grep -i "step"
done
timeout 300 python tools/k_sweep.py --ks 1,2,3,5,8 2>&1 | tail -12 | cut -c1-70 | tee gpurun_out/r2j_ksweep.txt
echo "== no unit deps"; IDQN_NO_UNIT_DEPS=1 timeout 300 python tools/k_sweep.py --ks 1,2,3,5,8 2>&1 | tail -12 | cut -c1-70
CTL_ROWS=12 timeout 300 python tools/cta_timeline.py 5 1 2 3 8 10 2>&1 > gpurun_out/r2_cta_timeline_deps_k5.txt
