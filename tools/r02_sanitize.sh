#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/r02_compute_sanitizer_$tool.txt 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|error" $O/r02_compute_sanitizer_$tool.txt | tail -14
done
