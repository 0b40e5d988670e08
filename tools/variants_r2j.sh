lib=ethicaltrajectoryplanning_b200/build/variants/tma/libetp_tma.so
for v in default tma; do
  if [ $v = tma ]; then export ETP_B200_LIB=$lib; else unset ETP_B200_LIB; fi
  echo "== $v cfg3"; python tools/run_step.py --steps 4 2>&1 | grep -E "^step [23]|debug"
  echo "== $v far";  python tools/run_step.py --steps 3 --lateral-spread 40 2>&1 | grep -E "^step 2"
  echo "== $v cfg2"; python tools/run_step.py --steps 4 --workload cfg2 2>&1 | grep -E "^step 3"
done
echo "== parity tma"; ETP_B200_LIB=$lib python -m pytest tests/test_gpu_golden.py tests/test_gpu_selection.py tests/test_gpu_edge_shapes.py tests/test_gpu_workloads.py -m gpu -x -q 2>&1 | tail -3
unset ETP_B200_LIB
python tools/run_step.py --steps 2 > gpurun_out/r2j_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2j_tma_off python tools/run_step.py --steps 2 > gpurun_out/r2j_ncu1.log 2>&1
ETP_B200_LIB=$lib ncu --set full --clock-control none --import-source on -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2j_tma_on python tools/run_step.py --steps 2 > gpurun_out/r2j_ncu2.log 2>&1
python tools/run_step.py --steps 2 --lateral-spread 40 > gpurun_out/r2j_plain2.log 2>&1 && ncu --set full --clock-control none -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2j_far_tma_off python tools/run_step.py --steps 2 --lateral-spread 40 > gpurun_out/r2j_ncu3.log 2>&1
ETP_B200_LIB=$lib ncu --set full --clock-control none -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2j_far_tma_on python tools/run_step.py --steps 2 --lateral-spread 40 > gpurun_out/r2j_ncu4.log 2>&1
