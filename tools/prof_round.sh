# Round profile: launch list of the bench command + full captures of the dominant kernels (1 GPU).
TAG=${1:-r1d}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_suite_polar_launches.csv python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fusion_loss_bwd_kernel -s 15 -c 1 -o gpurun_out/${TAG}_zkernel python tools/quick_bench.py 8x3072x4096 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:moment_fwd_kernel -s 1 -c 1 -o gpurun_out/${TAG}_fwdkernel python tools/quick_bench.py 8x3072x4096 > /dev/null 2>&1
MMIF_SERIAL=1 ncu --set full --clock-control none --import-source on -k regex:moment_fwd_kernel -s 6 -c 1 -o gpurun_out/${TAG}_vif17 python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
MMIF_SERIAL=1 ncu --set full --clock-control none --import-source on -k regex:pixel_metrics_kernel -s 1 -c 1 -o gpurun_out/${TAG}_pixel python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
MMIF_SERIAL=1 ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 1 -c 1 -o gpurun_out/${TAG}_hist python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
