#!/bin/bash
# quick persistent-kernel iteration: model-level GPU tests, phase timeline, short bench (no CPU legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_layer8b.py -m gpu -q -x > gpurun_out/r2_pytest_quick.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_quick.txt
for ring in 1; do
  echo "== JL_PD_RING=$ring"
  JL_PD_RING=$ring JL_PD_TRACE=1 timeout 300 python tools/ptrace.py 2>&1 | tail -10
  JL_PD_RING=$ring timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --prefill-tokens 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['step_frac'])"
done
