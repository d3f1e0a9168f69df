#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list of an un-graphed bench run, ncu --set full of one step.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-run}
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 300 --warmup 20 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --flags 1 > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/${tag}_step -f \
   --launch-skip 32 --launch-count 16 python tools/one_step.py 5 3 1 > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
