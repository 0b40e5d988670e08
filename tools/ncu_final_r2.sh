set -x
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /tmp/b0.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /tmp/ncu1.log 2>&1
python tools/run_step.py --steps 2 > /tmp/r0.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:^score_kernel -s 1 -c 1 -o /tmp/cfg3 python tools/run_step.py --steps 2 > /tmp/ncu2.log 2>&1 && \
python tools/ncu_summary.py /tmp/cfg3.ncu-rep > gpurun_out/r2_final_score_kernel_cfg3.txt
ETP_NO_GRAPH=1 python tools/closed_loop_profile.py cfg1 20 > /tmp/c0.log 2>&1 && \
ETP_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_final_launches_cycle_cfg1.csv python tools/closed_loop_profile.py cfg1 20 > /tmp/ncu3.log 2>&1
ETP_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:^score_kernel -s 12 -c 1 -o /tmp/cyc python tools/closed_loop_profile.py cfg1 20 > /tmp/ncu4.log 2>&1 && \
python tools/ncu_summary.py /tmp/cyc.ncu-rep > gpurun_out/r2_final_score_kernel_cycle_cfg1_cluster.txt
ETP_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_final_launches_cycle_cfg2.csv python tools/closed_loop_profile.py cfg2 20 > /tmp/ncu5.log 2>&1
tail -3 /tmp/ncu2.log /tmp/ncu4.log
ls -la gpurun_out/
