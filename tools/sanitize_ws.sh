# compute-sanitizer over every kernel family with the warp-specialised loss kernel as the default: tools/sanitize_ws.sh TAG
TAG=${1:-r2b}
: > gpurun_out/${TAG}_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" >> gpurun_out/${TAG}_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_once.py 2>&1 | tail -6 >> gpurun_out/${TAG}_sanitizer.txt
done
cat gpurun_out/${TAG}_sanitizer.txt
