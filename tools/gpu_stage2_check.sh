#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/stage2.log
for t in "eps_matches or two_stage or engine_trajectory or uncached_step" "full_model"; do
  echo "=== $t" >> gpurun_out/stage2.log
  timeout 900 python -m pytest tests/test_stage2_gpu.py -q -m gpu -k "$t" --tb=short -s -p no:cacheprovider 2>&1 | tail -60 >> gpurun_out/stage2.log
done
grep -E "^===|passed|failed|error|rel-L2|trace" gpurun_out/stage2.log
