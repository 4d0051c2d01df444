#!/usr/bin/env bash
# builds a trace-instrumented copy of the library (never the product .so) and runs the GEMM timeline tool
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# GEMM_SRC=<file>: trace an experimental kernel source instead (copied over csrc/gemm_conv.cu in this scratch copy only)
[ -n "${GEMM_SRC:-}" ] && cp "${GEMM_SRC}" remote-sensing-vision-language-diffusion-model_b200/csrc/gemm_conv.cu
B200SR_EXTRA_FLAGS="-DB200SR_GEMM_TRACE" B200SR_OUT=$PWD/remote-sensing-vision-language-diffusion-model_b200/b200sr/libb200sr_trace.so B200SR_BUILD_DIR=/tmp/build_trace \
  bash remote-sensing-vision-language-diffusion-model_b200/csrc/build.sh > /dev/null 2>&1 || { echo "trace build failed"; exit 1; }
B200SR_LIB=libb200sr_trace.so python tools/gemm_trace.py 2>&1 | tee gpurun_out/gemm_trace${TRACE_TAG:-}.txt
rm -f remote-sensing-vision-language-diffusion-model_b200/b200sr/libb200sr_trace.so
