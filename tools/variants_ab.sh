#!/bin/bash
# A/B of build variants of the library (tools/build_variants.py name=-DFLAG,...) on the GPU box: the headline step, the
# far-obstacle regime and the 10^4-candidate step, device-timed by tools/run_step.py.
#   bash tools/variants_ab.sh default pure unroll        (logs of round 2: profiles/r2_variants_*.log)
for v in "$@"; do
  if [ "$v" = default ]; then unset ETP_B200_LIB; else export ETP_B200_LIB=$PWD/ethicaltrajectoryplanning_b200/build/variants/$v/libetp_$v.so; fi
  echo "== $v cfg3"; python tools/run_step.py --steps 4 2>&1 | grep -E "^step [23]"
  echo "== $v far";  python tools/run_step.py --steps 3 --lateral-spread 40 2>&1 | grep -E "^step 2"
  echo "== $v cfg2"; python tools/run_step.py --steps 4 --workload cfg2 2>&1 | grep -E "^step 3"
done
