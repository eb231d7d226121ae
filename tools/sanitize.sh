#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on the GPU box: gpurun -- 'bash tools/sanitize.sh').
# Each tool gets its own bounded run; the summaries land in gpurun_out/sanitize_<tool>.txt.
set -u
mkdir -p gpurun_out
# default target: tools/sanitize.py (one small pass over every kernel family); TARGET="python -m pytest ... -m gpu -x -q" for tests
TARGET=${TARGET:-"python tools/sanitize.py"}
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  out=gpurun_out/sanitize_${tool}.txt
  timeout ${LIMIT:-900} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      $TARGET > $out.full 2>&1
  echo "exit=$?" > $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|Race reported|hazard|sanitize pass done|Traceback" $out.full | sort | uniq -c | sort -rn | head -40 >> $out
  rm -f $out.full.tmp
  tail -c 20000 $out.full > $out.tail; rm -f $out.full
done
cat gpurun_out/sanitize_*.txt
