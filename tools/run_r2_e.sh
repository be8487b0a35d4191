#!/bin/bash
# round-2 GPU call: final single-GPU verification -- smoke, whole test suite, prefill rate, bench lines (default and the driver's
# command), and re-captures of the kernels that changed after the first profiling pass
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2_pytest_e.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_e.txt
tail -6 gpurun_out/r2_pytest_e.txt
timeout 600 python tools/prefill_bench.py --batch 512,2048 > gpurun_out/r2_prefill.txt 2>&1
tail -2 gpurun_out/r2_prefill.txt
timeout 1500 python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
echo "bench rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_driver_cmd.json 2> gpurun_out/r2_bench_driver_cmd.err
echo "bench (driver command) rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_e.json', 'gpurun_out/r2_bench_driver_cmd.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}, d['roofline']['frac'])
        print('  prefill', d['config'].get('prefill')); c3=d['config'].get('config3_q8_batch8'); c3 and c3.pop('workload',None); print('  config3', c3)
        p=d.get('parity'); p.pop('note',None); print('  parity', p)
    except Exception as e:
        print('bench parse failed', f, e)
PY
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm8_kernel -s 3340 -c 1 -f -o $O/r2_gemm8_gate python tools/gemv_bench.py --batch > $O/r2_gemm8_gate.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attention_kernel -s 20 -c 1 -f -o $O/r2_prefill_attn python tools/prefill_bench.py --batch 2048 --repeat 1 > $O/r2_prefill_attn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_prefill_launches.csv python tools/prefill_bench.py --batch 2048 --repeat 1 > $O/r2_prefill_ncu.log 2>&1
ls -la $O/r2_gemm8_gate.ncu-rep $O/r2_prefill_attn.ncu-rep
