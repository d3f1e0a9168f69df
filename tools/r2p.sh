#!/bin/bash
# round-2 profile pass (1 GPU): launch list of the DEFAULT (graph + PDL) bench path, ncu --set full of one graph-replayed K=5 step,
# ncu launch list of smoke(), compute-sanitizer memcheck / racecheck / synccheck of two K=5 steps
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-prioritized --no-other-configs > gpurun_out/${tag}_launches.log 2>&1
echo "launches rc=$? rows=$(grep -c gpu__time gpurun_out/${tag}_launches.csv)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_smoke_launches.csv \
   python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke_launches.log 2>&1
echo "smoke launches rc=$? rows=$(grep -c gpu__time gpurun_out/${tag}_smoke_launches.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/${tag}_step -f \
   --launch-skip 33 --launch-count 15 python tools/one_step.py 5 4 0 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${tag}_ncu.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/one_step.py 5 2 0 > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_sanitizer_$tool.log)"
done
