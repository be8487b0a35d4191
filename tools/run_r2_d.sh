#!/bin/bash
# round-2 GPU call: whole single-GPU test suite + the full bench line (decode, e2e, prefill, long-context point, config 3, parity)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2_pytest_d.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_d.txt
tail -8 gpurun_out/r2_pytest_d.txt
timeout 1500 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
    print(d['config'].get('prefill')); print(d['config'].get('decode_long_context')); print(d['config'].get('config3_q8_batch8'))
    print(d.get('parity')); print(d.get('cpu_baseline')); print(d['roofline']['step_frac'])
except Exception as e:
    print('bench parse failed', e)
PY
