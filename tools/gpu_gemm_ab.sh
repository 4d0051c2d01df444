#!/usr/bin/env bash
# A/B of the GEMM epilogue: correctness first, then per-op in-graph timing and the step, for
#   (a) B200SR_GEMM_EPI_TMA=0   register / LSU epilogue everywhere (round-1 behaviour)
#   (b) default                 TMA residual load + TMA store for one-tile-per-CTA GEMMs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  echo "=== $name"
  env "$@" timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -3
  env "$@" timeout 600 python tools/bench_graph_ops.py gemm 2>&1 | tail -14 | tee gpurun_out/ab_ops_$name.txt
  env "$@" timeout 600 python tools/bench_graph_ops.py fixed 2>&1 | tail -20 | tee gpurun_out/ab_fixed_$name.txt
  env "$@" timeout 600 python tools/bench_graph_ops.py xattn 2>&1 | grep -E "xattn" | tee gpurun_out/ab_xattn_$name.txt
  env "$@" timeout 900 python bench.py --steps 10 --no-tiled --no-batched --no-images --no-cpu-baseline --no-gpu-eager > gpurun_out/ab_bench_$name.json 2> gpurun_out/ab_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/ab_bench_{sys.argv[1]}.json"))
    print("   step ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
          "gemm", round(d["roofline"]["by_kind"]["gemm"]["ms"], 2), "conv", round(d["roofline"]["by_kind"]["conv3x3"]["ms"], 2))
except Exception as e:
    print("   bench failed", e)
PY
}
run epi_lsu B200SR_GEMM_EPI_TMA=0
run epi_tma B200SR_GEMM_EPI_TMA=1
timeout 900 python -m pytest tests/test_stage2_gpu.py tests/test_sr3_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -3
TRACE_TAG=_epi_tma bash tools/gpu_trace.sh > /dev/null; grep -A9 "launch 4" gpurun_out/gemm_trace_epi_tma.txt | head -24
