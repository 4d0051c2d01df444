#!/usr/bin/env bash
# correctness (ops + stage-2 parity) then bench; everything logged under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/ops.log; tail -3 gpurun_out/ops.log
timeout 900 python -m pytest tests/test_stage2_gpu.py -q -m gpu --tb=short -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed|Error|assert" | tail -12
timeout 1200 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('ms/step', d['ms_per_step'], 'steps/s', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'])
r = d['roofline']; print('dense achieved', r['achieved'], 'frac', r['frac'], 'step tflops', r['step']['tflops'], r['step']['frac_of_peak'])
for k, v in r['by_kind'].items(): print(k, v)
print('clocks', d['clocks'])
PY
