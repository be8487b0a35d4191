#!/bin/bash
# round-2 GPU call: batched-decode tests + config 3 A/B of kernel variants (K split, flat vs tiled attention, warps per CTA)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_layer8b.py -m gpu -q --timeout 300 > gpurun_out/r2_pytest_f.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_f.txt
tail -6 gpurun_out/r2_pytest_f.txt
: > gpurun_out/r2_config3_ab2.txt
for v in "16" "8"; do
  JL_G8_WARPS=$v timeout 300 python tools/config3_bench.py 2>/dev/null | tail -1 | sed "s/^{/{\"JL_G8_WARPS\": \"$v\", /" >> gpurun_out/r2_config3_ab2.txt
done
cut -c1-330 gpurun_out/r2_config3_ab2.txt
timeout 600 python tools/gemv_bench.py --batch > gpurun_out/r2_gemv_batch.txt 2>&1
grep -v "f32" gpurun_out/r2_gemv_batch.txt | grep "M=8" | grep 8B
