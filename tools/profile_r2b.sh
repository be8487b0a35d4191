#!/bin/bash
# Round-2 launch lists (the first pass capped at 400/700 launches and was used up by the weight quantiser and the prompt prefill):
# synthetic weights written directly (no quantiser kernels), 8-token prompt, larger caps.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefill-tokens 0 --no-roofline --no-config3 --weights direct --prompt 8"
export JL_PD_TOKENS=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_launches_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_graph.csv $B --no-persistent > $O/r2_launches_graph.log 2>&1
python - <<'PY'
import csv, collections, re
for f in ('gpurun_out/r2_launches_bench.csv', 'gpurun_out/r2_launches_graph.csv'):
    rows = list(csv.reader(l for l in open(f) if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name')
    c = collections.Counter(re.sub(r'[<(].*', '', r[ki]) for r in rows[1:] if r[mi] == 'gpu__time_duration.sum')
    print(f, dict(c))
PY
