#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_n34.py > gpurun_out/r2_sanitizer_n34_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_n34_$tool.log | tail -1)"
done
tail -3 gpurun_out/r2_sanitizer_n34_memcheck.log
