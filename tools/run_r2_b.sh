#!/bin/bash
# round-2 GPU call: persistent decode kernel bring-up (tests without -x, phase timeline, bench) + the probe
mkdir -p gpurun_out
timeout 150 tools/micro/persist_probe > gpurun_out/r2_persist_probe.txt 2>&1
echo "probe rc=$?" >> gpurun_out/r2_persist_probe.txt
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_gpu_b.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_b.txt
tail -25 gpurun_out/r2_pytest_gpu_b.txt
JL_PD_TRACE=1 timeout 300 python tools/ptrace.py > gpurun_out/r2_ptrace_a.txt 2>&1
cat gpurun_out/r2_ptrace_a.txt | tail -14
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}, d['config'].get('decode_mode'), d.get('parity'), d['roofline']['step_frac'], d['config'].get('decode_long_context'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/r2_bench_b.err
cat gpurun_out/r2_persist_probe.txt
