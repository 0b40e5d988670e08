for v in base2 erfc7 nores e7nr; do
  lib=ethicaltrajectoryplanning_b200/build/variants/$v/libetp_$v.so
  echo "== $v cfg3"; ETP_B200_LIB=$lib python tools/run_step.py --steps 4 2>&1 | grep -E "^step [23]|debug"
  echo "== $v far";  ETP_B200_LIB=$lib python tools/run_step.py --steps 3 --lateral-spread 40 2>&1 | grep -E "^step 2"
done
echo "== parity of e7nr"; ETP_B200_LIB=ethicaltrajectoryplanning_b200/build/variants/e7nr/libetp_e7nr.so python tools/parity_report.py 2>&1 | grep -E "worst|mismatch" | head -12
echo "== loop breakdown"; python tools/loop_breakdown.py 2>&1 | head -45
