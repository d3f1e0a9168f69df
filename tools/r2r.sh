#!/bin/bash
# round-2 closing pass (1 GPU): tests, smoke, bench line, timelines, then the profile pass of tools/r2p.sh plus
# (a) DRAM bytes per kernel of the graph-replayed step with a WARM L2 (--cache-control none, one metric pass: no replay), and
# (b) compute-sanitizer memcheck of impala steps (small net with 8 / 6-channel pools: fast and generic pool kernels, tcgen05 + CUDA-core convs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2r_smoke.log
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -2 gpurun_out/r2r_bench.err
for k in 1 5 8; do timeout 300 python tools/kernel_timeline.py $k > gpurun_out/r2r_timeline_k$k.txt 2>&1; tail -1 gpurun_out/r2r_timeline_k$k.txt; done
timeout 300 python tools/impala_profile.py 5 > gpurun_out/r2r_impala_profile_k5.txt 2>&1; head -1 gpurun_out/r2r_impala_profile_k5.txt
timeout 300 python tools/impala_profile.py 1 > gpurun_out/r2r_impala_profile_k1.txt 2>&1; head -1 gpurun_out/r2r_impala_profile_k1.txt
bash tools/r2p.sh r02
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv \
   --log-file gpurun_out/r02_dram_warm_l2.csv --launch-skip 33 --launch-count 15 python tools/one_step.py 5 4 0 > gpurun_out/r02_dram_warm_l2.log 2>&1
echo "warm-L2 dram pass rc=$? rows=$(grep -c dram__bytes gpurun_out/r02_dram_warm_l2.csv)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/impala_step.py 2 2 > gpurun_out/r02_sanitizer_memcheck_impala.log 2>&1
echo "sanitizer memcheck impala rc=$?: $(grep -E 'ERROR SUMMARY' gpurun_out/r02_sanitizer_memcheck_impala.log)"
