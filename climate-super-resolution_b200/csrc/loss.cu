// Pixel loss of the training step: L1Loss() for esrgan / MSELoss() for srcnn, mean over ALL N*H*W pixels, un-masked
// (climsr/core/task.py:141; climsr/task/pl_generator_pre_training.py:29-30; climsr/task/pl_gan.py:41).
//
// One HBM-bound pass reads sr and hr once (8 B / pixel, 16-byte vector loads) and writes BOTH the loss value and
// d loss / d sr (4 B / pixel), so autograd's backward of the loss is a multiply by the upstream scalar instead of a
// second pass.  Deterministic: per-block partial sums in double, finalised by one block in a fixed order.
#include "loss.cuh"

namespace csr {

namespace {

constexpr int kLossThreads = 256;

template <int MODE>   // 0 = L1, 1 = MSE
__global__ void __launch_bounds__(kLossThreads)
pixel_loss_kernel(const float* __restrict__ sr, const float* __restrict__ hr, float* __restrict__ grad, long n, float inv_n,
                  double* __restrict__ partial) {
  double acc = 0.0;
  const long n4 = n >> 2;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(sr)[i];
    const float4 b = reinterpret_cast<const float4*>(hr)[i];
    const float d[4] = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
    float g[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (MODE == 0) {
        s += fabsf(d[k]);
        g[k] = d[k] > 0.f ? inv_n : (d[k] < 0.f ? -inv_n : 0.f);      // torch: sign(0) = 0
      } else {
        s += d[k] * d[k];
        g[k] = 2.f * d[k] * inv_n;
      }
    }
    acc += s;
    if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
  // ragged tail (n not a multiple of 4)
  for (long i = (n4 << 2) + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const float d = sr[i] - hr[i];
    if (MODE == 0) {
      acc += fabsf(d);
      if (grad) grad[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
    } else {
      acc += d * d;
      if (grad) grad[i] = 2.f * d * inv_n;
    }
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double warp_part[kLossThreads / 32];
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) t += warp_part[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void pixel_loss_finalize(const double* __restrict__ partial, int nblocks, double inv_n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nblocks; ++i) t += partial[i];
    *out = static_cast<float>(t * inv_n);
  }
}

}  // namespace

int pixel_loss_blocks(long n) {
  long b = (n / 4 + kLossThreads - 1) / kLossThreads;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

cudaError_t launch_pixel_loss(int mode, const float* sr, const float* hr, float* grad, long n, float* out, double* partial, cudaStream_t s) {
  const int nb = pixel_loss_blocks(n);
  const float inv_n = static_cast<float>(1.0 / static_cast<double>(n));
  if (mode == 0)
    pixel_loss_kernel<0><<<nb, kLossThreads, 0, s>>>(sr, hr, grad, n, inv_n, partial);
  else
    pixel_loss_kernel<1><<<nb, kLossThreads, 0, s>>>(sr, hr, grad, n, inv_n, partial);
  pixel_loss_finalize<<<1, 32, 0, s>>>(partial, nb, 1.0 / static_cast<double>(n), out);
  return cudaGetLastError();
}

}  // namespace csr
