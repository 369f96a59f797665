#!/bin/bash
# compute-sanitizer over tools/sanitizer_run.py (every kernel family once, small shapes).  Run on a GPU box:
#   gpurun --timeout 2400 -- 'bash tools/sanitize.sh gpurun_out'
# Writes <out>/sanitizer_{plain,memcheck,racecheck,synccheck,initcheck}.log; copy the summaries into profiles/.
out=${1:-gpurun_out}
mkdir -p "$out"
python tools/sanitizer_run.py > "$out/sanitizer_plain.log" 2>&1 || { echo "plain run failed"; tail -20 "$out/sanitizer_plain.log"; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check full"
  [ "$tool" = initcheck ] && extra="--track-unused-memory no"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 50 --target-processes all \
      python tools/sanitizer_run.py > "$out/sanitizer_$tool.log" 2>&1
  echo "$tool exit=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY|sanitizer workload ok' "$out/sanitizer_$tool.log" | tr '\n' ' ')"
done
