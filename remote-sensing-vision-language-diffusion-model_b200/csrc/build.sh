#!/usr/bin/env bash
# Build libb200sr.so (sm_100a only) in-tree next to the Python package.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${B200SR_OUT:-${HERE}/../b200sr/libb200sr.so}"   # B200SR_OUT: A/B builds (select with B200SR_LIB)
BUILD="${B200SR_BUILD_DIR:-${HERE}/build}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
# B200SR_EXTRA_FLAGS: debug builds only (e.g. -DB200SR_ATT_TRACE for tools/attn_trace.py)
read -r -a EXTRA <<< "${B200SR_EXTRA_FLAGS:-}"
FLAGS+=("${EXTRA[@]}")
mkdir -p "${BUILD}"
pids=()
for f in gemm_conv attention norm elementwise image capi; do
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${BUILD}/${f}.o" 2> "${BUILD}/${f}.log" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "${BUILD}"/*.log >&2; exit 1; fi
"${NVCC}" -shared -o "${OUT}" "${BUILD}"/{gemm_conv,attention,norm,elementwise,image,capi}.o -lcudart
echo "built ${OUT}"
