# Round-2 profile (1 GPU): launch list of the bench command, full capture of the dominant kernel AS LAUNCHED BY THE
# DROP-IN MODULES, the forward kernel, the suite launch list, compute-sanitizer over every kernel family.
TAG=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_bench_launches.csv > gpurun_out/${TAG}_bench_launch_list.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_suite_polar_launches.csv python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_suite_polar_launches.csv 2 > gpurun_out/${TAG}_suite_polar_launch_list.txt
ncu --set full --clock-control none --import-source on -k regex:fusion_loss_bwd_kernel -s 2 -c 1 -o gpurun_out/${TAG}_zkernel python tools/modules_once.py 8x3072x4096 4 > gpurun_out/${TAG}_zkernel_run.log 2>&1
bash tools/ncu_summary.sh gpurun_out/${TAG}_zkernel.ncu-rep > gpurun_out/${TAG}_zkernel.txt
python tools/ncu_phases.py gpurun_out/${TAG}_zkernel.ncu-rep >> gpurun_out/${TAG}_zkernel.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:moment_fwd_kernel -s 1 -c 1 -o gpurun_out/${TAG}_fwdkernel python tools/quick_bench.py 8x3072x4096 > /dev/null 2>&1
bash tools/ncu_summary.sh gpurun_out/${TAG}_fwdkernel.ncu-rep > gpurun_out/${TAG}_fwdkernel.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" >> gpurun_out/${TAG}_sanitizer.txt
  compute-sanitizer --tool $tool python tools/sanitize_once.py 2>&1 | tail -4 >> gpurun_out/${TAG}_sanitizer.txt
done
rm -f gpurun_out/${TAG}_fwdkernel.ncu-rep
ls -la gpurun_out/${TAG}_*
