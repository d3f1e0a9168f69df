#!/bin/bash
mkdir -p gpurun_out
for tag in fwd1 fwd2 dgrad1 dgrad2 wgrad0; do
  IDQN_TL=$tag timeout 120 python tools/timeline.py 5 > gpurun_out/r2g_tl_$tag.txt 2>&1
  echo "$tag: $(wc -l < gpurun_out/r2g_tl_$tag.txt) events, last: $(tail -1 gpurun_out/r2g_tl_$tag.txt)"
done
