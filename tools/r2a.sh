#!/bin/bash
# round-2 GPU pass A: K sweep, ncu on smoke() with and without graph / PDL, compute-sanitizer on one K=5 step
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))"
timeout 600 python tools/k_sweep.py --ks 1,2,3,5,8 --steps 300 > gpurun_out/r2a_ksweep.jsonl 2> gpurun_out/r2a_ksweep.err
cat gpurun_out/r2a_ksweep.jsonl
for f in 0 16 1 17; do
  IDQN_FLAGS=$f timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_ncu_smoke_f$f.csv \
     python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_ncu_smoke_f$f.log 2>&1
  echo "ncu smoke flags=$f rc=$? launches=$(grep -c gpu__time gpurun_out/r2a_ncu_smoke_f$f.csv)"
  tail -3 gpurun_out/r2a_ncu_smoke_f$f.log
done
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/one_step.py 5 2 0 > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; tail -4 gpurun_out/r2a_sanitizer_$tool.log
done
