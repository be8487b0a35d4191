for cfg in 0 1 2 3; do for ah in 0 3 6; do
  psm=3; if [ $cfg -ge 2 ]; then psm=2; fi
  JL_GEMV_CFG=$cfg JL_GEMV_AHEAD=$ah JL_GEMV_PER_SM=$psm timeout 120 python tools/gemv_bench.py --quick 2>&1 | grep "^\["
done; done
