run() { echo "=== $*"; cp jlama_b200/variants/$1.so jlama_b200/libjlama_b200.so; timeout 150 python tools/ktrace.py 2>&1 | grep -A14 "^# llama\|^# averages" | grep -v "^# kind" ; }
cp jlama_b200/libjlama_b200.so /tmp/cur.so
run w3
run w4
cp /tmp/cur.so jlama_b200/libjlama_b200.so
