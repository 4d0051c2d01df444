#!/usr/bin/env bash
# round-2 GPU pass: full GPU test suite, smoke, default bench line; logs under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -q -m gpu --tb=short -s -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/gputests.log 2>&1
grep -E "rel-L2|PSNR|trace|passed|failed|Error|error|assert|xfail|XFAIL" gpurun_out/gputests.log | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ -z "${SKIP_BENCH:-}" ]; then
timeout ${BENCH_TIMEOUT:-1500} python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench.json'))
except Exception as e:
    print("bench.json unreadable:", e); raise SystemExit(0)
print('ms/step', d['ms_per_step'], 'steps/s', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches_per_step'])
r = d['roofline']; print('dense achieved', r['achieved'], 'frac', r['frac'], 'share', r['share_of_step_time'], 'eager ms', r['eager_step_ms'])
for k, v in r['by_kind'].items(): print(' ', k, v)
print('clocks', d['clocks'])
for k in ('cpu_baseline', 'gpu_eager_baseline', 'batched', 'tiled_x8', 'images_per_s'):
    print(k, json.dumps(d.get(k)))
PY
fi
