set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_suite_polar_launches.csv python tools/suite_once.py 32 1024 1224 2 > gpurun_out/s2_suite_polar.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_suite_tno_launches.csv python tools/suite_once.py 21 480 640 2 > gpurun_out/s2_suite_tno.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fusion_loss_bwd_kernel -s 1 -c 1 -o gpurun_out/s2_zkernel python tools/quick_bench.py 8x3072x4096 > gpurun_out/s2_zk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:moment_fwd_kernel -s 1 -c 1 -o gpurun_out/s2_fwdkernel python tools/quick_bench.py 8x3072x4096 > gpurun_out/s2_fk.log 2>&1
python tools/quick_bench.py > gpurun_out/s2_quick.log 2>&1
tail -5 gpurun_out/s2_quick.log
