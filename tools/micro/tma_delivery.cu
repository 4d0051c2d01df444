// What caps operand delivery L2 -> SM?  Every CTA streams 32 KB "stages" (two 128-row x 64-column bf16 boxes)
// from an L2-resident matrix in a deep TMA ring, no MMA.  Variant 0: each CTA loads its own boxes.  Variant 1:
// clusters of 2, each CTA loads ONE box and multicasts it to both CTAs (same bytes land in every SM, half the L2
// reads).  If variant 1 delivers more bytes per second per SM, the cap is on the L2 side and multicast helps GEMMs.
// Build (from csrc/): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. ../../tools/micro/tma_delivery.cu -o ../../tools/micro/tma_delivery -L../b200sr -lb200sr -Xlinker -rpath -Xlinker '$ORIGIN/../../remote-sensing-vision-language-diffusion-model_b200/b200sr' -lcuda
#include "common.cuh"
#include <cstdio>
#include <vector>
using namespace b200sr;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void remote_arrive(uint64_t* bar, uint32_t rank) {
  uint32_t addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}

constexpr int STAGES = 6, BOX_BYTES = 128 * 64 * 2;

template <int kMulticast>
__global__ void __launch_bounds__(64) delivery_kernel(const __grid_constant__ CUtensorMap tm, int iters, int rows_total, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const uint32_t rank = kMulticast ? ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMulticast ? 2 : 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (kMulticast) cluster_sync();
  const int cl = kMulticast ? blockIdx.x / 2 : blockIdx.x;
  long long t0 = clock64();
  if (threadIdx.x == 0) {  // producer
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect_tx(&full[st], 2 * BOX_BYTES);
      uint8_t* dst = smem + st * 2 * BOX_BYTES;
      const int k0 = (it * 64) % 4096;
      const int r0 = ((cl * 2) * 128) % rows_total, r1 = ((cl * 2 + 1) * 128) % rows_total;
      if (!kMulticast) {
        tma_load_2d(dst, &tm, &full[st], k0, r0);
        tma_load_2d(dst + BOX_BYTES, &tm, &full[st], k0, r1);
      } else {
        // this CTA loads box `rank` and multicasts it into both CTAs of the cluster
        tma_load_2d_mc(dst + rank * BOX_BYTES, &tm, &full[st], k0, rank ? r1 : r0, 3);
      }
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer: just frees the stage (in both CTAs when multicasting)
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[st], ph);
      if (!kMulticast) mbar_arrive(&empty[st]);
      else { remote_arrive(&empty[st], 0); remote_arrive(&empty[st], 1); }
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (kMulticast) cluster_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
  const int rows = 148 * 256, cols = 4096;  // 310 MB matrix would miss L2: use rows that wrap -> 2 x 148 boxes x 4096 cols = 310 MB? keep it L2 resident:
  const int rows_total = 4096;              // 4096 x 4096 bf16 = 32 MB, L2 resident
  __nv_bfloat16* d; cudaMalloc(&d, (size_t)rows_total * cols * 2); cudaMemset(d, 0, (size_t)rows_total * cols * 2);
  (void)rows;
  CUtensorMap tm;
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows_total}; uint64_t strides[1] = {(uint64_t)cols * 2}; uint32_t box[2] = {64, 128};
  if (make_tmap_bf16(&tm, d, 2, dims, strides, box)) { printf("tmap failed\n"); return 1; }
  long long* dout; cudaMalloc(&dout, 8);
  const size_t smem = STAGES * 2 * BOX_BYTES + 1024;
  cudaFuncSetAttribute(delivery_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(delivery_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4000;
  for (int v = 0; v < 2; ++v) {
    long long h = 0; float ms = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      cudaError_t e = v == 0 ? launch_k(delivery_kernel<0>, dim3(148), dim3(64), smem, 0, 1, tm, iters, rows_total, dout)
                             : launch_k(delivery_kernel<1>, dim3(148), dim3(64), smem, 0, 2, tm, iters, rows_total, dout);
      cudaEventRecord(e1);
      cudaError_t e2 = cudaDeviceSynchronize();
      if (e != cudaSuccess || e2 != cudaSuccess) { printf("error %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
      cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
    const double bytes_per_sm = (double)iters * 2 * BOX_BYTES;
    printf("%s: %.3f ms, delivered %.2f TB/s chip-wide (%.1f GB/s per SM, %.0f cycles per 32 KB stage), L2 reads %.2f TB/s\n",
           v == 0 ? "own loads        " : "pairwise multicast", ms, bytes_per_sm * 148 / ms / 1e9, bytes_per_sm / ms / 1e6,
           (double)h / iters, bytes_per_sm * 148 / ms / 1e9 / (v == 0 ? 1 : 2));
  }
  return 0;
}
