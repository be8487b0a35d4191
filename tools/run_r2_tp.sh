#!/bin/bash
# tensor- / expert-parallel run on N GPUs of one box (N = $1): the world-size-N parity tests (torchrun inside pytest), the bench at
# N ranks, and at N = 8 the Mixtral-8x7B expert-parallel line
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_gpu_tp.py -m gpu -q -s --timeout 600 -k "[$N-" > gpurun_out/r2_pytest_tp$N.txt 2>&1
echo "pytest tp rc=$?"
grep -E "tp=|passed|failed|skipped|Error|error" gpurun_out/r2_pytest_tp$N.txt | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29511 + N)) bench.py --gpus $N --steps 128 --warmup 16 > gpurun_out/r2_bench_tp$N.json 2> gpurun_out/r2_bench_tp$N.err
echo "bench tp$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_tp$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','gpu_launches')}, d['config'].get('decode_mode'), d.get('parity'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/r2_bench_tp$N.err
if [ $N -eq 8 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --model mixtral-8x7b --weights direct --steps 64 --warmup 8 > gpurun_out/r2_bench_mixtral_tp$N.json 2> gpurun_out/r2_bench_mixtral_tp$N.err
  echo "bench mixtral tp$N rc=$?"
  cat gpurun_out/r2_bench_mixtral_tp$N.json | cut -c1-1500
  tail -4 gpurun_out/r2_bench_mixtral_tp$N.err
fi
