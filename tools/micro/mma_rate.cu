// Microbenchmark: cycles per tcgen05.mma for the shapes the attention kernel issues.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I<csrc> tools/micro/mma_rate.cu -o tools/micro/mma_rate -lcuda
#include "common.cuh"
#include <cstdio>
using namespace b200sr;

// mode 0: SS  M128 N128 K16 (QK^T), K-major A and B, 4 MMAs per "block"
// mode 1: TS  M128 N64  K16 (PV), A in TMEM, B MN-major, 8 MMAs per block
// mode 2: TS  M128 N64  K16, B K-major descriptor (for comparison only; data meaningless)
// mode 3: SS  M128 N64  K16, B MN-major
// mode 4: alternate mode 0 block and mode 1 block (what the kernel does)
// mode 5: SS  M128 N256 K16
__global__ void __launch_bounds__(128, 2) mma_rate_kernel(int mode, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc_s = umma_idesc_bf16_f32(128, 128, 0, 0);
    const uint32_t idesc_o = umma_idesc_bf16_f32(128, 64, 0, 1);
    const uint32_t idesc_ok = umma_idesc_bf16_f32(128, 64, 0, 0);
    const uint32_t idesc_256 = umma_idesc_bf16_f32(128, 256, 0, 0);
    const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
    const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(smem + 32768), 1024, 1024);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 0 || mode == 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, qdesc + 2 * k, kdesc + 2 * k, idesc_s, 1);
      }
      if (mode == 1 || mode == 4) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ts(tmem + 192, tmem + 128 + 8 * k, vdesc + (2048 >> 4) * k, idesc_o, 1);
      }
      if (mode == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ts(tmem + 192, tmem + 128 + 8 * k, kdesc + 2 * (k & 3), idesc_ok, 1);
      }
      if (mode == 3) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tmem + 192, qdesc + 2 * (k & 3), vdesc + (2048 >> 4) * k, idesc_o, 1);
      }
      if (mode == 5) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, qdesc + 2 * k, kdesc + 2 * k, idesc_256, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const size_t smem = 64 * 1024 + 1024 + 16384 * 0;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"SS M128 N128 (QK, 4/blk)", "TS M128 N64 B=MN-major (PV, 8/blk)", "TS M128 N64 B=K-major", "SS M128 N64 B=MN-major", "QK+PV alternating", "SS M128 N256 (4/blk)"};
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
    for (int mode = 0; mode < 6; ++mode) {
      const int reps = 2000;
      long long h = 0;
      for (int it = 0; it < 2; ++it) {
        mma_rate_kernel<<<148 * ctas_per_sm, 128, smem>>>(mode, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("ctas/sm %d  %-38s  %8.1f cycles per block (per CTA)\n", ctas_per_sm, names[mode], double(h) / reps);
    }
  return 0;
}
