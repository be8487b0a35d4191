run() { echo "=== $* $KT_ARGS"; env "$@" timeout 150 python tools/ktrace.py $KT_ARGS 2>&1 | grep -A14 "^# llama\|^# averages" | grep -v "^# kind" ; }
run JL_PDL=1
