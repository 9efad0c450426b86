# metric-suite check: parity tests, per-kernel launch list (cold-cache, serialised) and event-timed suite throughput
python -m pytest tests/test_metric_gpu.py -x -q 2>&1 | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/suite_polar_launches.csv python tools/suite_once.py 32 1024 1224 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/suite_tno_launches.csv python tools/suite_once.py 21 480 640 2 > /dev/null 2>&1
python tools/suite_time.py
