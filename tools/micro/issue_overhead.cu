// What does the per-K-chunk bookkeeping of the GEMM mainloop cost the MMA issue thread?
// chunk = 4 x tcgen05.mma (M128, N=160 or 256, K16, SS).  Variants add, per chunk: a commit, a
// tcgen05.fence::after_thread_sync, and an mbarrier wait that is already satisfied.
#include "common.cuh"
#include <cstdio>
using namespace b200sr;

__global__ void __launch_bounds__(128, 1) issue_kernel(int variant, int n, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[8];
  __shared__ uint64_t done, ready;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(&done, 1);
    mbar_init(&ready, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, n, 0, 0);
    const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
    long long t0 = clock64();
    if (variant <= 3) {
      for (int r = 0; r < reps; ++r) {
        if (variant >= 3) mbar_wait(&ready, 1);          // parity of the phase BEFORE the first completion: returns at once
        if (variant >= 2) tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
        if (variant >= 1) umma_commit(&bars[r & 7]);
      }
    } else if (variant == 5 || variant == 6) {
      // 2 (variant 5) or 3 (variant 6) chunks per barrier wait / commit
      const int per = variant == 5 ? 2 : 3;
      for (int r = 0; r < reps; r += per) {
        mbar_wait(&ready, 1);
        tc_fence_after();
        for (int j = 0; j < per; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem, adesc + 2 * k + 128 * j, bdesc + 2 * k + 64 * j, idesc, 1);
        }
        umma_commit(&bars[r & 7]);
      }
    } else {
      // software-pipelined probe: the next chunk's barrier is probed BEFORE this chunk's MMAs are issued and the
      // predicate is only consumed afterwards, so the TRYWAIT latency hides under the MMA issue
      bool ok = mbar_try_wait(&ready, 1);
      for (int r = 0; r < reps; ++r) {
        if (!ok) mbar_wait(&ready, 1);
        tc_fence_after();
        const bool ok_next = mbar_try_wait(&ready, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
        umma_commit(&bars[r & 7]);
        ok = ok_next;
      }
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const size_t smem = 64 * 1024 + 1024;
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"4 MMAs back to back", "+ commit per chunk", "+ fence::after_thread_sync", "+ satisfied mbarrier wait",
                         "wait probed one chunk ahead", "2 chunks per wait + commit", "3 chunks per wait + commit"};
  for (int n : {160, 256})
    for (int v = 0; v < 7; ++v) {
      long long h = 0;
      for (int it = 0; it < 2; ++it) {
        issue_kernel<<<148, 128, smem>>>(v, n, 1998, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("N %3d  %-32s %7.1f cycles per chunk (floor %d)\n", n, names[v], double(h) / 1998, 4 * 128 * n / 256);
    }
  return 0;
}
