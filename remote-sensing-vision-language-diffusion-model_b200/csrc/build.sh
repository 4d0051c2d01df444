#!/usr/bin/env bash
# Build libb200sr.so (sm_100a only) in-tree next to the Python package.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../b200sr/libb200sr.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
# B200SR_EXTRA_FLAGS: debug builds only (e.g. -DB200SR_ATT_TRACE for tools/attn_trace.py)
read -r -a EXTRA <<< "${B200SR_EXTRA_FLAGS:-}"
FLAGS+=("${EXTRA[@]}")
mkdir -p "${HERE}/build"
pids=()
for f in gemm_conv attention norm elementwise image capi; do
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/build/${f}.o" 2> "${HERE}/build/${f}.log" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "${HERE}"/build/*.log >&2; exit 1; fi
"${NVCC}" -shared -o "${OUT}" "${HERE}"/build/{gemm_conv,attention,norm,elementwise,image,capi}.o -lcudart
echo "built ${OUT}"
