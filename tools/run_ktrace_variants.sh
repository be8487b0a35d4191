run() { echo "=== $* $KT_ARGS"; env "$@" timeout 150 python tools/ktrace.py $KT_ARGS 2>&1 | grep -A14 "^# llama\|^# averages" | grep -v "^# kind" ; }
run JL_X=0
KT_ARGS="--flags 16" run JL_X=pdl
