#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): launch lists + full captures of the dominant kernels.
# Numbers printed by processes running under ncu are never bench values.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prefill-tokens 0 --no-roofline --no-config3"
export JL_PD_TOKENS=1   # one token per cooperative launch, so that a captured launch is exactly one decode step
# 1. launch list of the bench command (persistent decode kernel: one launch per token)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_launches_bench.log 2>&1
# 2. full capture of one persistent decode step (the whole token: all weight GEMVs, attention, lm_head, argmax)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pdecode_kernel -s 6 -c 1 -f -o $O/r2_pdecode $B > $O/r2_pdecode.log 2>&1
# 3. the per-op graph path: launch list of one step + the lm_head GEMV (generic kernel, F32 activations x Q4, 128256 rows)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_graph.csv $B --no-persistent > $O/r2_launches_graph.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -c 2 -f -o $O/r2_lm_head $B --no-persistent > $O/r2_lm_head.log 2>&1
# 4. batched-decode GEMM (8 sessions) and the prefill kernels
# (gemv_bench --batch launches 208 gemm8 kernels per (shape, M, mode): 3340 = gate M=8 q8, 4590 = down M=8 q8)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm8_kernel -s 3340 -c 1 -f -o $O/r2_gemm8_gate python tools/gemv_bench.py --batch > $O/r2_gemm8_gate.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm8_kernel -s 4590 -c 1 -f -o $O/r2_gemm8_down python tools/gemv_bench.py --batch > $O/r2_gemm8_down.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_q4_tc_kernel -s 40 -c 1 -f -o $O/r2_gemm_tc python tools/prefill_bench.py --batch 2048 --repeat 1 > $O/r2_gemm_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prefill_attention_kernel -s 20 -c 1 -f -o $O/r2_prefill_attn python tools/prefill_bench.py --batch 2048 --repeat 1 > $O/r2_prefill_attn.log 2>&1
ls -la $O/*.ncu-rep
