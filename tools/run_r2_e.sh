#!/bin/bash
# round-2 GPU call: smoke, whole single-GPU test suite, prefill rate, launch lists, the full bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2_pytest_e.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_e.txt
tail -6 gpurun_out/r2_pytest_e.txt
timeout 600 python tools/prefill_bench.py --batch 512,2048 > gpurun_out/r2_prefill.txt 2>&1
cat gpurun_out/r2_prefill.txt | tail -3
bash tools/profile_r2b.sh > gpurun_out/r2_profile_b.log 2>&1
tail -3 gpurun_out/r2_profile_b.log | cut -c1-600
timeout 1500 python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_e.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
    print(d['config'].get('prefill')); print(d['config'].get('config3_q8_batch8'))
    p=d.get('parity'); p.pop('note',None); print(p); print(d.get('cpu_baseline')); print(d['roofline'])
except Exception as e:
    print('bench parse failed', e)
PY
