#!/bin/bash
# Round-1 profiling pass (run on the GPU box through gpurun): launch list of the bench command + full captures.
# Numbers printed by processes running under ncu are never bench values.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefill-tokens 0 --no-roofline"
# 1. launch list of the bench command (every kernel node of the replayed graph, serialised, cold cache)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_r1.csv $B > $O/launches_r1.log 2>&1
# 2. full capture of five consecutive decode GEMV launches (qkv, o_proj, gate/up, down, qkv) and one attention
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_decode_kernel -s 300 -c 5 -f -o $O/gemv_decode_r1 $B > $O/gemv_full_r1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_decode_attention -s 60 -c 1 -f -o $O/attention_r1 $B > $O/attn_full_r1.log 2>&1
# 3. the lm_head GEMV (generic kernel, F32 activations)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 3 -c 1 -f -o $O/lm_head_r1 $B > $O/lmhead_full_r1.log 2>&1
# 4. tcgen05 prefill GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_q4_tc_kernel -s 30 -c 1 -f -o $O/gemm_tc_r1 python tools/gemm_tc_bench.py > $O/gemm_tc_full_r1.log 2>&1
ls -la $O/*.ncu-rep
