#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/k_sweep.py --ks 1,3,5,8 2>&1 | tail -4 | cut -c1-70; }
timeout 900 python -m pytest tests/test_gpu_learn.py -x -q -m gpu 2>&1 | tail -3
run A=1
run IDQN_NO_WG_AFTER=1
run IDQN_WG_OVERLAP=80
run IDQN_WG_OVERLAP=110
timeout 200 python tools/kernel_timeline.py 5 2>&1 | tail -17
