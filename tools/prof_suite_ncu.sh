for k in hist_kernel pixel_metrics_kernel "moment_fwd_kernel<17" halve_kernel; do
  nm=$(echo $k | tr -dc 'a-z0-9_')
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -o gpurun_out/s2_$nm python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
