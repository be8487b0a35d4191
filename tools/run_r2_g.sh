#!/bin/bash
# round-2 GPU call: Q8_0 weights on the tensor-core prefill path (tests) + config 3 with it + the driver's own bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "tensor_core or prefill" > gpurun_out/r2_pytest_g.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_g.txt
tail -5 gpurun_out/r2_pytest_g.txt
timeout 400 python tools/config3_bench.py 2>/dev/null | tail -1 | cut -c1-700
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_driver_cmd.json 2> gpurun_out/r2_bench_driver_cmd.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_driver_cmd.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}, d['roofline']['frac'])
    print(d['config'].get('config3_q8_batch8'))
    p=d.get('parity'); p.pop('note',None); print(p)
except Exception as e:
    print('bench parse failed', e)
PY
