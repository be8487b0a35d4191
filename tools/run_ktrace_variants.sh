run() { echo "=== $*"; cp jlama_b200/variants/$1.so jlama_b200/libjlama_b200.so; timeout 150 python tools/ktrace.py 2>&1 | grep -A14 "^# llama\|^# averages" | grep -v "^# kind" ; }
cp jlama_b200/libjlama_b200.so /tmp/cur.so
run v3
run v0
cp /tmp/cur.so jlama_b200/libjlama_b200.so
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -2
