lib=ethicaltrajectoryplanning_b200/build/variants/tma/libetp_tma.so
cap() {  # name, env lib or "", extra args
  name=$1; shift; L=$1; shift
  if [ -n "$L" ]; then export ETP_B200_LIB=$L; else unset ETP_B200_LIB; fi
  ncu --set full --clock-control none -k regex:^score_kernel -s 1 -c 1 -o /tmp/$name python tools/run_step.py --steps 2 "$@" > gpurun_out/r2l_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/r2l_$name.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r2l_$name.raw.csv 2>/dev/null
  unset ETP_B200_LIB
}
python tools/run_step.py --steps 2 > gpurun_out/r2l_plain.log 2>&1 && cap tma_off ""
cap tma_on $lib
python tools/run_step.py --steps 2 --lateral-spread 40 > gpurun_out/r2l_plain2.log 2>&1 && cap far_tma_off "" --lateral-spread 40
cap far_tma_on $lib --lateral-spread 40
python tools/run_step.py --steps 2 --workload cfg2 > gpurun_out/r2l_plain3.log 2>&1 && cap cfg2_tma_off "" --workload cfg2
cap cfg2_tma_on $lib --workload cfg2
du -sh gpurun_out
