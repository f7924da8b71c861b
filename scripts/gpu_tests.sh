#!/bin/bash
# Runs every -m gpu test file in its own process (a trapped kernel poisons only its own context);
# logs land in gpurun_out/ which gpurun merges back.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
TAG=${1:-t}
shift
FILES=${@:-conv ops nms model fullsize teacher_forced}
for t in $FILES; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -rA --tb=short -s > gpurun_out/${TAG}_$t.log 2>&1
  echo "$t exit $?"
  grep -E "passed|failed|error" gpurun_out/${TAG}_$t.log | tail -3
done
