#!/bin/bash
# copy the artefacts of the last tools/r2p.sh pass from gpurun_out/ into profiles/ (traffic JSON stamped with the source hash,
# ncu step summary with its header, launch lists, sanitizer logs, SASS summary)
set -e
python tools/ncu_traffic.py gpurun_out/r02_step.ncu-rep > profiles/ncu_traffic.json
python tools/ncu_summary.py gpurun_out/r02_step.ncu-rep > /tmp/ncu_new.md
python - <<'PY'
import re
new = open('/tmp/ncu_new.md').read()
body = new[new.index('| # | kernel'):]
rows = [l for l in body.splitlines() if re.match(r"\| \d+ \|", l)]
tot = sum(float(re.match(r"\| \d+ \| [^|]+ \| ([0-9.]+) \|", l).group(1)) for l in rows)
upd = [l for l in rows if "dense_wgrad_adam_kernel" in l][0].split("|")
hdr = f"""# ncu --set full --clock-control none --import-source on, one GRAPH-REPLAYED K=5 step (tools/one_step.py 5 4 0, default flags: CUDA graph + programmatic dependent launch), round 2 final sources

15 kernel nodes of one step as ncu replays them one by one (cold caches, serialised: compare shares, not absolutes).  The window starts in the backward pass of one step and ends in the backward pass of the next.  Grids are those of the graph: the Dense_0 update (`dense_wgrad_adam_kernel`) on {int(float(upd[4]))} CTAs and the two conv data-gradient kernels (`conv_taps_kernel<1, 2>`) on the other SMs -- in the graph these run side by side on three branches, here one after the other.  DRAM bytes per launch feed `profiles/ncu_traffic.json`.

"""
open('profiles/r02_ncu_step_summary.md', 'w').write(hdr + body + f"\nSum of the 15 kernel durations (serialised, cold): {tot:.1f} us; dense_wgrad_adam_kernel {float(upd[3]):.1f} us = {100 * float(upd[3]) / tot:.1f} % of it\n")
print("ncu sum", tot, "update", upd[3], upd[4])
PY
for f in launches.csv smoke_launches.csv sanitizer_memcheck.log sanitizer_racecheck.log sanitizer_synccheck.log; do cp gpurun_out/r02_$f profiles/r02_$f; done
python tools/sass_summary.py > profiles/r02_sass_summary.txt 2>&1
[ -f gpurun_out/r2_cta_timeline_k5.txt ] && cp gpurun_out/r2_cta_timeline_k5.txt profiles/r02_cta_timeline_k5.txt
python -c "
import json; from bench import source_sha
print('traffic sha', json.load(open('profiles/ncu_traffic.json'))['source_sha'], 'sources', source_sha())"
