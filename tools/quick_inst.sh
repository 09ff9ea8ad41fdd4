#!/bin/bash
# instruction count + time of the extraction kernels for one 64-frame batch (run on the GPU box)
ncu --clock-control none --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:"k_fast_blur|k_describe|k_octree|k_resize_tma" --launch-skip 10 -c 10 --csv python tools/profile_batch.py 64 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(l for l in sys.stdin if l.startswith('\"')))
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['Kernel Name'][:24], d['Metric Name'], d['Metric Value'])
"
