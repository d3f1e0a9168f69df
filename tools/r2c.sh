#!/bin/bash
# round-2 GPU pass C (1 GPU): full GPU suite, CTA-0 pipeline timelines of the conv kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2c_pytest.log
for tag in fwd0 fwd1 fwd2 dgrad1 dgrad2 wgrad0 wgrad1 wgrad2; do
  IDQN_TL=$tag timeout 120 python tools/timeline.py 5 > gpurun_out/r2c_tl_$tag.txt 2>&1
  echo "$tag: $(wc -l < gpurun_out/r2c_tl_$tag.txt) events, last: $(tail -1 gpurun_out/r2c_tl_$tag.txt)"
done
