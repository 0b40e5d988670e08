for v in base pipe sig estrin all3; do
  lib=ethicaltrajectoryplanning_b200/build/variants/$v/libetp_$v.so
  echo "== $v cfg3"; ETP_B200_LIB=$lib python tools/run_step.py --steps 4 2>&1 | grep -E "^step [23]|debug"
  echo "== $v far";  ETP_B200_LIB=$lib python tools/run_step.py --steps 3 --lateral-spread 40 2>&1 | grep -E "^step 2"
  echo "== $v cfg2"; ETP_B200_LIB=$lib python tools/run_step.py --steps 4 --workload cfg2 2>&1 | grep -E "^step 3"
done
