#!/bin/bash
# round-2 GPU pass B: full GPU test suite, free-running drift report, ncu on smoke() (default graph path), in-graph timelines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2b_pytest.log
timeout 600 python tests/test_gpu_golden.py mlp_k3 cnn_k1 cnn_k3 cnn_k5 cnn_k8 > gpurun_out/r2b_drift.log 2>&1; tail -5 gpurun_out/r2b_drift.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_ncu_smoke.csv \
   python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_ncu_smoke.log 2>&1
echo "ncu smoke rc=$? launches=$(grep -c gpu__time gpurun_out/r2b_ncu_smoke.csv)"; tail -2 gpurun_out/r2b_ncu_smoke.log
for k in 1 5 8; do timeout 300 python tools/kernel_timeline.py $k > gpurun_out/r2b_timeline_k$k.txt 2>&1; cat gpurun_out/r2b_timeline_k$k.txt; done
timeout 600 python tools/k_sweep.py --ks 1,5,8 --steps 300 > gpurun_out/r2b_ksweep.jsonl 2> gpurun_out/r2b_ksweep.err; cut -c1-120 gpurun_out/r2b_ksweep.jsonl
