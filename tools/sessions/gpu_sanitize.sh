#!/usr/bin/env bash
# Dev: compute-sanitizer memcheck over tools/sanitize_small.py, one kernel family at a time (bounded).
mkdir -p gpurun_out
for part in render guided refine split mlp; do
  echo "== $part"
  timeout -s KILL ${SAN_TIMEOUT:-150} compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 7 \
      python tools/sanitize_small.py $part > gpurun_out/sanitize_$part.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|ok|Invalid|out of bounds|Error" gpurun_out/sanitize_$part.log | head -8
done
