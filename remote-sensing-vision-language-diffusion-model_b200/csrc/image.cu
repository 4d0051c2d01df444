// b200sr — output side of the restoration pipeline: wavelet colour fix and the
// bicubic resize + uint8 pack of Tensor2PIL.  fp32 NCHW images, memory-bound elementwise / stencil kernels.
//
// Reference semantics:
//   wavelet_blur / wavelet_decomposition / wavelet_reconstruction     utils/colorfix.py:73-119
//     blur = depthwise 3x3 [1 2 1; 2 4 2; 1 2 1] / 16 with dilation r = 2^level, replicate padding r;
//     high += image - low; image = low (5 levels); result = content_high + style_low
//   Tensor2PIL                                                         models/util.py:159-166
//     F.interpolate(size = (h0, w0), mode = "bicubic") (align_corners False, A = -0.75, clamped taps),
//     x * 127.5 + 127.5, clip to [0, 255], truncate to uint8, HWC
#include "common.cuh"

namespace b200sr {

__global__ void wavelet_level_kernel(const float* __restrict__ img, float* __restrict__ low, float* __restrict__ high,
                                     int first, int H, int W, int r, size_t total) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const float* plane = img + (i / (static_cast<size_t>(W) * H)) * static_cast<size_t>(W) * H;
    const int ym = max(y - r, 0), yp = min(y + r, H - 1), xm = max(x - r, 0), xp = min(x + r, W - 1);
    const float* r0 = plane + static_cast<size_t>(ym) * W;
    const float* r1 = plane + static_cast<size_t>(y) * W;
    const float* r2 = plane + static_cast<size_t>(yp) * W;
    // same tap order as a row-major 3x3 kernel
    float l = 0.0625f * r0[xm];
    l = fmaf(0.125f, r0[x], l);
    l = fmaf(0.0625f, r0[xp], l);
    l = fmaf(0.125f, r1[xm], l);
    l = fmaf(0.25f, r1[x], l);
    l = fmaf(0.125f, r1[xp], l);
    l = fmaf(0.0625f, r2[xm], l);
    l = fmaf(0.125f, r2[x], l);
    l = fmaf(0.0625f, r2[xp], l);
    low[i] = l;
    if (high != nullptr) high[i] = (first ? 0.f : high[i]) + (r1[x] - l);
  }
}
int wavelet_level(const float* img, float* low, float* high, int first, int BC, int H, int W, int radius,
                  cudaStream_t stream) {
  if (BC <= 0 || H <= 0 || W <= 0 || radius <= 0 || img == low) return B200SR_EINVAL;
  const size_t total = static_cast<size_t>(BC) * H * W;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  launch_k(wavelet_level_kernel, dim3(grid), dim3(256), 0, stream, 1, img, low, high, first, H, W, radius, total);
  return cudaGetLastError() == cudaSuccess ? B200SR_OK : B200SR_ELAUNCH;
}

__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                               size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = a[i] + b[i];
}
int add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t stream) {
  if (n <= 0) return B200SR_EINVAL;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  launch_k(add_f32_kernel, dim3(grid), dim3(256), 0, stream, 1, a, b, out, static_cast<size_t>(n));
  return cudaGetLastError() == cudaSuccess ? B200SR_OK : B200SR_ELAUNCH;
}

// cubic convolution coefficients of torch's upsample_bicubic2d (A = -0.75)
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__global__ void image_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int C, int H, int W, int OH,
                                   int OW, float sh, float sw) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(OH) * OW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % OW), oy = static_cast<int>(i / OW);
    const float ry = sh * (oy + 0.5f) - 0.5f, rx = sw * (ox + 0.5f) - 0.5f;
    const int iy = static_cast<int>(floorf(ry)), ix = static_cast<int>(floorf(rx));
    float cy[4], cx[4];
    cubic_coeffs(ry - iy, cy);
    cubic_coeffs(rx - ix, cx);
    for (int c = 0; c < C; ++c) {
      const float* plane = x + static_cast<size_t>(c) * H * W;
      float v = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float* row = plane + static_cast<size_t>(min(max(iy - 1 + a, 0), H - 1)) * W;
        float rv = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) rv += cx[b] * row[min(max(ix - 1 + b, 0), W - 1)];
        v += cy[a] * rv;
      }
      v = fminf(fmaxf(v * 127.5f + 127.5f, 0.f), 255.f);
      out[i * C + c] = static_cast<uint8_t>(v);   // truncation, as numpy's astype(uint8) on a clipped float
    }
  }
}
int image_to_u8(const float* x, void* out, int C, int H, int W, int OH, int OW, cudaStream_t stream) {
  if (C <= 0 || C > 4 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) return B200SR_EINVAL;
  const size_t total = static_cast<size_t>(OH) * OW;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  launch_k(image_to_u8_kernel, dim3(grid), dim3(256), 0, stream, 1, x, reinterpret_cast<uint8_t*>(out), C, H, W, OH, OW,
           static_cast<float>(H) / OH, static_cast<float>(W) / OW);
  return cudaGetLastError() == cudaSuccess ? B200SR_OK : B200SR_ELAUNCH;
}

}  // namespace b200sr
