run() { echo "=== $*"; env "$@" timeout 150 python tools/ktrace.py 2>&1 | grep -A14 "^# llama\|^# averages" | grep -v "^# kind" ; }
run JL_X=0
