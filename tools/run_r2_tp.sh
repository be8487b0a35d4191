#!/bin/bash
# tensor-parallel bring-up on N GPUs of one box: parity tests (torchrun inside pytest), then bench at 2 ranks
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_gpu_tp.py -m gpu -q -s > gpurun_out/r2_pytest_tp.txt 2>&1
echo "pytest tp rc=$?"
grep -E "tp=|passed|failed|skipped|Error|error" gpurun_out/r2_pytest_tp.txt | tail -20
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 64 --warmup 8 > gpurun_out/r2_bench_tp$N.json 2> gpurun_out/r2_bench_tp$N.err
echo "bench tp$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_tp$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','gpu_launches')}, d['config'].get('decode_mode'), d.get('parity'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -6 gpurun_out/r2_bench_tp$N.err
