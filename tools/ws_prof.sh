# ncu capture of the warp-specialised kernel + per-region stall table: tools/ws_prof.sh TAG
TAG=${1:-ws}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fusion_loss_ws_kernel -s 1 -c 1 -o gpurun_out/${TAG} python tools/ws_ab.py 8x3072x4096 > gpurun_out/${TAG}_run.log 2>&1
bash tools/ncu_summary.sh gpurun_out/${TAG}.ncu-rep > gpurun_out/${TAG}.txt
python tools/ws_regions.py gpurun_out/${TAG}.ncu-rep >> gpurun_out/${TAG}.txt
cat gpurun_out/${TAG}.txt
