for v in default pure unroll both; do
  if [ $v = default ]; then unset ETP_B200_LIB; else export ETP_B200_LIB=ethicaltrajectoryplanning_b200/build/variants/$v/libetp_$v.so; fi
  echo "== $v cfg3"; python tools/run_step.py --steps 4 2>&1 | grep -E "^step [23]"
  echo "== $v far";  python tools/run_step.py --steps 3 --lateral-spread 40 2>&1 | grep -E "^step 2"
  echo "== $v cfg2"; python tools/run_step.py --steps 4 --workload cfg2 2>&1 | grep -E "^step 3"
done
