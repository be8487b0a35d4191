#!/bin/bash
# round-2 first GPU call: full GPU test suite, the persistent-kernel probe, one bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1
nproc >> gpurun_out/r2_gpus.txt
timeout 300 tools/micro/persist_probe > gpurun_out/r2_persist_probe.txt 2>&1
echo "probe rc=$?" >> gpurun_out/r2_persist_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.txt
tail -5 gpurun_out/r2_pytest_gpu.txt
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench rc=$?"
cat gpurun_out/r2_bench_a.json
tail -3 gpurun_out/r2_persist_probe.txt
