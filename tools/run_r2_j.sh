#!/bin/bash
# round-2 GPU call: sanity of the last build (smoke + model / GPT-2 tests) and the long-context decode points
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_gpt2.py tests/test_gpu_checkpoint.py -m gpu -q --timeout 300 2>&1 | tail -3
timeout 900 python tools/long_context_bench.py > gpurun_out/r2_long_context.txt 2>gpurun_out/r2_long_context.err
cat gpurun_out/r2_long_context.txt; tail -3 gpurun_out/r2_long_context.err
