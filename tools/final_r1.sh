set -x
O=gpurun_out
timeout 900 python bench.py > $O/bench_final_r1.json 2> $O/bench_final_r1.err; tail -2 $O/bench_final_r1.err
timeout 600 python bench.py --impl reference --steps 32 --warmup 3 > $O/bench_ref_r1.json 2> $O/bench_ref_r1.err; tail -2 $O/bench_ref_r1.err
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -2
