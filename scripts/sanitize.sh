#!/bin/bash
# compute-sanitizer over one small step of every workload (run on the GPU box through gpurun):
#   bash scripts/sanitize.sh [tag]     -> gpurun_out/<tag>_sanitizer_{memcheck,racecheck}.txt
tag=${1:-r02}
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_step.py \
      > gpurun_out/${tag}_sanitizer_$tool.txt 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_step: done" gpurun_out/${tag}_sanitizer_$tool.txt
done
