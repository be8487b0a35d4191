#!/bin/bash
# round-2 GPU call: GPT-2 family (BASELINE config 1 as a parity case) + regression of the model tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gpt2.py tests/test_gpu_model.py -m gpu -q --timeout 300 > gpurun_out/r2_pytest_i.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_i.txt
tail -30 gpurun_out/r2_pytest_i.txt | cut -c1-300
