#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report of one K=5 step: DRAM bytes (read + written) per launch of the
Dense_0 kernels, stamped with the hash of the kernel sources they were captured from (bench.py refuses a stale file).

    python tools/ncu_traffic.py gpurun_out/r02_step.ncu-rep > profiles/ncu_traffic.json"""
import csv
import io
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import source_sha

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tags = {"dense_wgrad_adam_kernel": "dense_wgrad_adam_L3", "dense_stream_kernel<0>": "dense_fwd_L3", "dense_stream_kernel<1>": "dense_dgrad_L3"}
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of one graph-replayed "
                   "K=5 step (batch 32) summarised in profiles/r02_ncu_step_summary.md; read by bench.py for roofline.traffic",
       "source_sha": source_sha(), "step_total_bytes": 0}
for r in data:
    name = r[idx["Kernel Name"]]
    b = sum(float(r[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out["step_total_bytes"] += int(b)
    for k, tag in tags.items():
        if k in name:
            out[tag] = int(b)
print(json.dumps(out, indent=1))
