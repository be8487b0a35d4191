#!/bin/bash
# round-2 GPU call: 2..8-row int8 tensor-core GEMM (jl_gemm8.cu) + tiled prefill attention -- tests, per-shape timing, prefill rate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q --timeout 180 > gpurun_out/r2_pytest_c.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_c.txt
tail -25 gpurun_out/r2_pytest_c.txt
timeout 600 python tools/gemv_bench.py --batch > gpurun_out/r2_gemv_batch.txt 2>&1
grep -v "f32" gpurun_out/r2_gemv_batch.txt | grep "M=[18]" | grep 8B | tail -40
grep "f32" gpurun_out/r2_gemv_batch.txt | grep "lm_head"
timeout 600 python tools/prefill_bench.py > gpurun_out/r2_prefill.txt 2>&1
cat gpurun_out/r2_prefill.txt | tail
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prefill_launches.csv python tools/prefill_bench.py --batch 2048 --repeat 1 > gpurun_out/r2_prefill_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open('gpurun_out/r2_prefill_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
n = len(rows) - 1
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1 + n // 2:]:   # second half = the timed pass (the warm-up pass has the same launches)
    v = float(r[vi].replace(',', '')); v = v / 1000.0 if r[ui] == 'ns' else v
    k = re.sub(r'<.*', '', r[ki]); tot[k] += v; cnt[k] += 1
s = sum(tot.values())
for k, v in tot.most_common(12):
    print('%-44s %6d launches %10.1f us  %5.1f%%' % (k, cnt[k], v, 100 * v / s))
print('total %.1f us' % s)
PY
