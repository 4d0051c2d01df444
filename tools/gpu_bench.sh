#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python -m pytest tests/test_stage2_gpu.py -q -m gpu -k "two_stage or engine_trajectory" --tb=short -p no:cacheprovider 2>&1 | tail -5
