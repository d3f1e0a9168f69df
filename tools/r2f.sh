#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_learn.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2f_pytest.log
timeout 600 python tools/k_sweep.py --ks 1,5 --steps 300 > gpurun_out/r2f_ksweep.jsonl 2> gpurun_out/r2f_ksweep.err; cut -c1-100 gpurun_out/r2f_ksweep.jsonl; tail -3 gpurun_out/r2f_ksweep.err
IDQN_FLAGS=2048 timeout 600 python tools/k_sweep.py --ks 5 --steps 300 2>/dev/null | cut -c1-100
timeout 300 python tools/kernel_timeline.py 5 > gpurun_out/r2f_timeline_k5.txt 2>&1; cat gpurun_out/r2f_timeline_k5.txt
timeout 300 python tools/kernel_timeline.py 1 2>&1 | tail -16
