# Round-2b profile (1 GPU): launch list of the bench command and a full capture of the dominant kernel — the warp-specialised
# fusion_loss_ws_kernel AS LAUNCHED BY THE DROP-IN MODULES — with its per-role block table.
TAG=${1:-r2b}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_bench_launches.csv > gpurun_out/${TAG}_bench_launch_list.txt
ncu --set full --clock-control none --import-source on -k regex:fusion_loss_ws_kernel -s 2 -c 1 -o gpurun_out/${TAG}_wskernel python tools/modules_once.py 8x3072x4096 4 > gpurun_out/${TAG}_wskernel_run.log 2>&1
bash tools/ncu_summary.sh gpurun_out/${TAG}_wskernel.ncu-rep > gpurun_out/${TAG}_wskernel.txt
python tools/ws_blocks.py gpurun_out/${TAG}_wskernel.ncu-rep >> gpurun_out/${TAG}_wskernel.txt 2>/dev/null
ls -la gpurun_out/${TAG}_*
