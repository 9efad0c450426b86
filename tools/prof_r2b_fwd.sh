# Round-2b captures of the forward strip kernel and of one suite call (after the predicated sums / slot pointers)
TAG=${1:-r2b}
ncu --set full --clock-control none --import-source on -k regex:moment_fwd_kernel -s 1 -c 1 -o gpurun_out/${TAG}_fwdkernel python tools/quick_bench.py 8x3072x4096 > /dev/null 2>&1
bash tools/ncu_summary.sh gpurun_out/${TAG}_fwdkernel.ncu-rep > gpurun_out/${TAG}_fwdkernel.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_suite_polar_launches.csv python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_suite_polar_launches.csv 2 > gpurun_out/${TAG}_suite_polar_launch_list.txt
rm -f gpurun_out/${TAG}_fwdkernel.ncu-rep
