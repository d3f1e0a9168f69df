#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/k_sweep.py --ks 1,3,5,8 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['K'], d['ms_per_step'], 'update alone', d['kernel_us'].get('dense_wgrad_adam_L3'))"; }
timeout 900 python -m pytest tests/test_gpu_learn.py -x -q -m gpu 2>&1 | tail -3
run A=1
run IDQN_WG_OVERLAP=0
run IDQN_WG_OVERLAP=64
run IDQN_WG_OVERLAP=72
run IDQN_WG_OVERLAP=80
run IDQN_WG_OVERLAP=88
run IDQN_WG_OVERLAP=104
