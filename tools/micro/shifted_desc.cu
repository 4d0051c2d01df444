// Feasibility test for a halo-tile 3x3 convolution: can a tcgen05.mma A operand (K-major, 128B swizzle) be a
// SHIFTED window of a larger swizzled shared-memory tile, with 8-row groups 10 pixel rows (1280 B) apart?
// The halo tile holds 18 x 10 pixels x 64 channels, written the way TMA would write one box into a 1024-aligned
// buffer (16-byte chunk c of row p stored at chunk c ^ (p & 7)).  For tap (kh, kw) the A operand is the 16 x 8
// pixel window starting at pixel (kh, kw): row m = dy * 8 + dx -> halo row (dy + kh) * 10 + dx + kw.
// D[m, n] = sum_c A[m, c] * B[n, c] with B = a known matrix; compared with a CPU reference.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace b200sr;

__global__ void __launch_bounds__(128) shifted_kernel(const __nv_bfloat16* halo /*[180][64]*/, const __nv_bfloat16* bmat /*[64][64]*/,
                                                      float* out /*[128][64]*/, int shift_rows, int base_off_mode, int sbo_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 180 rows x 128 B = 23040 B -> pad to 24 KB
  uint8_t* sB = smem + 24576;    // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // emulate TMA SWIZZLE_128B writes
  for (int i = threadIdx.x; i < 180 * 8; i += blockDim.x) {
    const int p = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + p * 128 + ((c ^ (p & 7)) << 4)) = reinterpret_cast<const uint4*>(halo)[p * 8 + c];
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int p = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + p * 128 + ((c ^ (p & 7)) << 4)) = reinterpret_cast<const uint4*>(bmat)[p * 8 + c];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, 64, 0, 0);
    const uint32_t a_addr = smem_u32(sA) + shift_rows * 128;
    uint64_t adesc = umma_smem_desc_sw128(a_addr, 16, sbo_bytes);
    if (base_off_mode == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sB), 16, 1024);
    for (int k = 0; k < 4; ++k) umma_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<float> halo(180 * 64), bm(64 * 64);
  srand(1);
  for (auto& v : halo) v = bf((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : bm) v = bf((rand() % 2001 - 1000) / 1000.0f);
  std::vector<__nv_bfloat16> hh(halo.size()), hb(bm.size());
  for (size_t i = 0; i < halo.size(); ++i) hh[i] = __float2bfloat16(halo[i]);
  for (size_t i = 0; i < bm.size(); ++i) hb[i] = __float2bfloat16(bm[i]);
  __nv_bfloat16 *dh, *db; float* dout;
  cudaMalloc(&dh, hh.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dh, hh.data(), hh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 24576 + 8192 + 1024;
  cudaFuncSetAttribute(shifted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> out(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int shift = kh * 10 + kw;
        shifted_kernel<<<1, 128, smem>>>(dh, db, dout, shift, mode, 1280);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) {
          const int dy = m >> 3, dx = m & 7, row = (dy + kh) * 10 + dx + kw;
          for (int n = 0; n < 64; ++n) {
            double acc = 0;
            for (int c = 0; c < 64; ++c) acc += (double)halo[row * 64 + c] * bm[n * 64 + c];
            maxerr = fmax(maxerr, fabs(acc - out[m * 64 + n]));
          }
        }
        printf("base_offset_mode %d tap (%d,%d) shift %2d rows: max abs err %.4g %s\n", mode, kh, kw, shift, maxerr, maxerr < 1e-2 ? "OK" : "WRONG");
      }
  return 0;
}
