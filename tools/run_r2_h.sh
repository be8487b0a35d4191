#!/bin/bash
# round-2 GPU call: fused BF16 producers + 8-warp prefill attention -- tests and prefill rate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_checkpoint.py -m gpu -q --timeout 300 > gpurun_out/r2_pytest_h.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_h.txt
tail -5 gpurun_out/r2_pytest_h.txt
timeout 600 python tools/prefill_bench.py --batch 512,2048 > gpurun_out/r2_prefill.txt 2>&1
tail -3 gpurun_out/r2_prefill.txt
JL_PA_WARPS=4 timeout 600 python tools/prefill_bench.py --batch 2048 2>&1 | tail -1
