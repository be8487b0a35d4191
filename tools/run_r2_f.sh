#!/bin/bash
# round-2 GPU call: new tests (batched long-context decode, KV persistence, sparse operands) + config 3 A/B of kernel variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q --timeout 300 > gpurun_out/r2_pytest_f.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_f.txt
tail -6 gpurun_out/r2_pytest_f.txt
: > gpurun_out/r2_config3_ab.txt
for v in "0 1" "8 1" "0 0" "1 1" "4 1"; do
  set -- $v
  JL_G8_KSPLIT=$1 JL_ATTN_FLAT=$2 timeout 300 python tools/config3_bench.py 2>/dev/null | tail -1 >> gpurun_out/r2_config3_ab.txt
done
cut -c1-420 gpurun_out/r2_config3_ab.txt
timeout 600 python tools/gemv_bench.py --batch > gpurun_out/r2_gemv_batch.txt 2>&1
grep -v "f32" gpurun_out/r2_gemv_batch.txt | grep "M=8" | grep 8B
