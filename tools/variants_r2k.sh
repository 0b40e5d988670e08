lib=ethicaltrajectoryplanning_b200/build/variants/tma/libetp_tma.so
python tools/run_step.py --steps 2 > gpurun_out/r2k_plain.log 2>&1 && ncu --set full --clock-control none -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2k_tma_off python tools/run_step.py --steps 2 > gpurun_out/r2k_ncu1.log 2>&1
ETP_B200_LIB=$lib ncu --set full --clock-control none -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2k_tma_on python tools/run_step.py --steps 2 > gpurun_out/r2k_ncu2.log 2>&1
python tools/run_step.py --steps 2 --lateral-spread 40 > gpurun_out/r2k_plain2.log 2>&1 && ncu --set full --clock-control none -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2k_far_tma_off python tools/run_step.py --steps 2 --lateral-spread 40 > gpurun_out/r2k_ncu3.log 2>&1
ETP_B200_LIB=$lib ncu --set full --clock-control none -k regex:^score_kernel -s 1 -c 1 -o gpurun_out/r2k_far_tma_on python tools/run_step.py --steps 2 --lateral-spread 40 > gpurun_out/r2k_ncu4.log 2>&1
ls -la gpurun_out
