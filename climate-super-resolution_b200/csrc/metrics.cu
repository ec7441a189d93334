// Masked validation/test step of the reference as two HBM-bound passes (+ two tiny finalisers).
//
// Reference arithmetic (climsr/core/task.py:262-300, 342-380; climsr/metrics/regression_accuracy.py:15-22;
// climsr/data/normalization.py:63-84,115): denormalise sr per sample, zero-fill sr/hr/denorm_sr/original where
// mask == 0, then L1/MSE loss and 16 metrics, every mean taken over ALL pixels.  The reference runs ~10 full-tensor
// passes plus 16 torchmetrics calls; here pass 1 reads sr, hr, original, mask exactly once (16 B / pixel) and
// produces every sum / count / min / max, pass 2 reads sr, hr, mask once more (12 B / pixel) for the 11x11
// Gaussian SSIM.  Reductions are deterministic: per-block partials in double, reduced in a fixed order.
#include "metrics.cuh"

#include <cfloat>

namespace csr {

namespace {

constexpr int kP1Threads = 256;
constexpr int kSums = 16;     // 8 accuracy counts, |d|, d^2, t, t^2, smape, mape, l1, mse(normalised)
constexpr int kMinMax = 6;    // min/max of original, masked sr, masked hr
constexpr int kPart = kSums + kMinMax;
constexpr float kMapeEps = 1.17e-06f;

// Per-thread accumulators.  T = float for the z-score scaler (the reference's StandardScaler keeps float32:
// ``arr * std + mean`` with Python-float constants, normalization.py:115) and double for the min-max scaler: there the
// per-sample ``min`` / ``max`` reach MinMaxScaler._denormalize as float64 (N,) tensors (pandas columns through
// default_collate), so ``arr.permute(1,2,3,0) - min_`` promotes the denormalised tensor to float64 and every metric on it
// - the eight |p - t| <= eps counts in particular - is evaluated in float64 (normalization.py:70-82, regression_accuracy.py:18).
template <typename T>
struct Acc {
  float cnt[8];
  T s[5];            // sum|d|, sum d^2, sum t, sum t^2, smape terms           (denormalised pair)
  float sn[3];       // mape terms, sum|dn|, sum dn^2                            (normalised pair, always float32)
  float mn[3], mx[3];
};

template <typename T>
__device__ __forceinline__ void acc_pixel(Acc<T>& a, float sr, float hr, float orig, float mask, T dscale, T dmin, float zmean, float zstd) {
  const bool land = mask != 0.f;                        // (~mask.bool()) -> 0.0, task.py:288-291
  T den;
  if constexpr (sizeof(T) == 4) den = __fadd_rn(__fmul_rn(sr, zstd), zmean);          // two roundings, like the two torch ops (no FMA)
  else den = (static_cast<double>(sr) - dmin) / dscale;                               // out = (arr - min_) / scale: a true division
  const float srn = land ? sr : 0.f;
  const float hrn = land ? hr : 0.f;
  const T dsr = land ? den : static_cast<T>(0);
  const T org = land ? static_cast<T>(orig) : static_cast<T>(0);
  const T d = fabs(dsr - org);
  a.cnt[0] += d <= static_cast<T>(0.1);  a.cnt[1] += d <= static_cast<T>(0.25); a.cnt[2] += d <= static_cast<T>(0.5);
  a.cnt[3] += d <= static_cast<T>(0.75); a.cnt[4] += d <= static_cast<T>(1.0);  a.cnt[5] += d <= static_cast<T>(1.25);
  a.cnt[6] += d <= static_cast<T>(1.5);  a.cnt[7] += d <= static_cast<T>(2.0);
  a.s[0] += d;
  a.s[1] += d * d;
  a.s[2] += org;
  a.s[3] += org * org;
  if constexpr (sizeof(T) == 4) a.s[4] += __fdividef(2.f * d, fmaxf(fabsf(org) + fabsf(dsr), kMapeEps));   // fast division: 2 ulp, metrics are held to 1e-4
  else a.s[4] += static_cast<double>(__fdividef(2.f * static_cast<float>(d), fmaxf(static_cast<float>(fabs(org) + fabs(dsr)), kMapeEps)));
  const float dn = fabsf(srn - hrn);
  a.sn[0] += __fdividef(dn, fmaxf(fabsf(hrn), kMapeEps));
  a.sn[1] += dn;
  a.sn[2] += dn * dn;
  const float orgf = static_cast<float>(org);
  a.mn[0] = fminf(a.mn[0], orgf); a.mx[0] = fmaxf(a.mx[0], orgf);
  a.mn[1] = fminf(a.mn[1], srn); a.mx[1] = fmaxf(a.mx[1], srn);
  a.mn[2] = fminf(a.mn[2], hrn); a.mx[2] = fmaxf(a.mx[2], hrn);
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// grid = (blocks_per_image, N).  Each thread walks groups of 4 pixels of one image.
template <typename T>
__global__ void __launch_bounds__(kP1Threads)
metrics_pass1_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ orig,
                     const float* __restrict__ mask, const double* __restrict__ mn, const double* __restrict__ mx, float zmean, float zstd,
                     double ra, double rb, double eps, long hw, double* __restrict__ partials) {
  const int n = blockIdx.y;
  T dmin = 0, dscale = 1;
  if constexpr (sizeof(T) == 8) {
    // scale = (b-a)/((max-min)+eps); min_ = a - min*scale; out = (arr - min_)/scale   (normalization.py:70-82), all float64
    dscale = (rb - ra) / ((mx[n] - mn[n]) + eps);
    dmin = ra - mn[n] * dscale;
  }
  Acc<T> a;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.cnt[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) a.s[i] = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) { a.sn[i] = 0.f; a.mn[i] = FLT_MAX; a.mx[i] = -FLT_MAX; }
  const long base = static_cast<long>(n) * hw;
  const long groups = (hw + 3) >> 2;
  const bool vec = (hw & 3) == 0;
  for (long g = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; g < groups; g += static_cast<long>(gridDim.x) * blockDim.x) {
    const long p = base + g * 4;
    if (vec) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(sr + p));
      const float4 h4 = __ldg(reinterpret_cast<const float4*>(hr + p));
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(orig + p));
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(mask + p));
      acc_pixel<T>(a, s4.x, h4.x, o4.x, m4.x, dscale, dmin, zmean, zstd);
      acc_pixel<T>(a, s4.y, h4.y, o4.y, m4.y, dscale, dmin, zmean, zstd);
      acc_pixel<T>(a, s4.z, h4.z, o4.z, m4.z, dscale, dmin, zmean, zstd);
      acc_pixel<T>(a, s4.w, h4.w, o4.w, m4.w, dscale, dmin, zmean, zstd);
    } else {
      for (int k = 0; k < 4 && g * 4 + k < hw; ++k)
        acc_pixel<T>(a, sr[p + k], hr[p + k], orig[p + k], mask[p + k], dscale, dmin, zmean, zstd);
    }
  }
  __shared__ double sh_s[kP1Threads / 32][kSums];
  __shared__ float sh_m[kP1Threads / 32][kMinMax];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kSums; ++i) {
    // partial layout (kSums): 8 counts, sum|d|, sum d^2, sum t, sum t^2, smape, mape, l1, mse(normalised)
    const double own = i < 8 ? static_cast<double>(a.cnt[i]) : i < 13 ? static_cast<double>(a.s[i - 8]) : static_cast<double>(a.sn[i - 13]);
    const double v = warp_sum(own);
    if (lane == 0) sh_s[warp][i] = v;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float lo = warp_min(a.mn[i]), hi = warp_max(a.mx[i]);
    if (lane == 0) { sh_m[warp][2 * i] = lo; sh_m[warp][2 * i + 1] = hi; }
  }
  __syncthreads();
  if (threadIdx.x < kPart) {
    double* out = partials + (static_cast<long>(blockIdx.y) * gridDim.x + blockIdx.x) * kPart;
    if (threadIdx.x < kSums) {
      double v = 0;
      for (int w = 0; w < kP1Threads / 32; ++w) v += sh_s[w][threadIdx.x];
      out[threadIdx.x] = v;
    } else {
      const int i = threadIdx.x - kSums;
      float v = sh_m[0][i];
      for (int w = 1; w < kP1Threads / 32; ++w) v = (i & 1) ? fmaxf(v, sh_m[w][i]) : fminf(v, sh_m[w][i]);
      out[threadIdx.x] = v;
    }
  }
}

// Single block of 256 threads: fixed-order (deterministic) tree reduction of the pass-1 partials -> totals[kPart] (double)
// and the SSIM constants c1, c2.  Thread t owns partial-block slice t, t+256, ... for every quantity; the tree then
// combines the 256 lane sums in a fixed pattern.
__global__ void __launch_bounds__(256)
metrics_reduce1_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ totals, float* __restrict__ c12) {
  __shared__ double sh[kPart][256];
  __shared__ double tot[kPart];
  const int t = threadIdx.x;
  for (int i = 0; i < kPart; ++i) {
    const bool is_sum = i < kSums, is_max = ((i - kSums) & 1) != 0;
    double v = is_sum ? 0.0 : (is_max ? -DBL_MAX : DBL_MAX);
    for (int bk = t; bk < nblocks; bk += 256) {
      const double q = partials[static_cast<long>(bk) * kPart + i];
      v = is_sum ? v + q : (is_max ? fmax(v, q) : fmin(v, q));
    }
    sh[i][t] = v;
  }
  __syncthreads();
  for (int stride = 128; stride > 0; stride >>= 1) {
    if (t < stride)
      for (int i = 0; i < kPart; ++i) {
        const bool is_sum = i < kSums, is_max = ((i - kSums) & 1) != 0;
        const double a = sh[i][t], q = sh[i][t + stride];
        sh[i][t] = is_sum ? a + q : (is_max ? fmax(a, q) : fmin(a, q));
      }
    __syncthreads();
  }
  if (t < kPart) {
    totals[t] = sh[t][0];
    tot[t] = sh[t][0];
  }
  __syncthreads();
  if (t == 0) {
    // torchmetrics SSIM: data_range = max(p.max()-p.min(), t.max()-t.min()); c = (k*range)^2
    const double range = fmax(tot[kSums + 3] - tot[kSums + 2], tot[kSums + 5] - tot[kSums + 4]);
    c12[0] = static_cast<float>((0.01 * range) * (0.01 * range));
    c12[1] = static_cast<float>((0.03 * range) * (0.03 * range));
  }
}

// ---- SSIM: 11x11 Gaussian (sigma 1.5), valid region only (torchmetrics reflect-pads by 5 and then crops 5, so
// padding never reaches a retained output).  One block = 32x32 outputs; separable filter through shared memory.
constexpr int kST = 32;            // output tile edge
constexpr int kSI = kST + 10;      // input tile edge
__constant__ float c_gauss[11];

__global__ void __launch_bounds__(256)
ssim_kernel(const float* __restrict__ sr, const float* __restrict__ hr, const float* __restrict__ mask, int H, int W,
            const float* __restrict__ c12, double* __restrict__ partials) {
  __shared__ float s_p[kSI][kSI + 1];
  __shared__ float s_t[kSI][kSI + 1];
  __shared__ float s_h[5][kSI][kST + 1];
  __shared__ double s_red[8];
  const int n = blockIdx.z;
  const int oy = 5 + blockIdx.y * kST, ox = 5 + blockIdx.x * kST;   // first output pixel of this tile
  const float* psr = sr + static_cast<long>(n) * H * W;
  const float* phr = hr + static_cast<long>(n) * H * W;
  const float* pm = mask + static_cast<long>(n) * H * W;
  for (int i = threadIdx.x; i < kSI * kSI; i += blockDim.x) {
    const int r = i / kSI, c = i - r * kSI;
    const int y = min(oy - 5 + r, H - 1), x = min(ox - 5 + c, W - 1);
    const long q = static_cast<long>(y) * W + x;
    const bool land = pm[q] != 0.f;
    s_p[r][c] = land ? psr[q] : 0.f;
    s_t[r][c] = land ? phr[q] : 0.f;
  }
  __syncthreads();
  // Horizontal pass with a register sliding window: one thread = 8 consecutive outputs of one row (18 loads of p and t
  // instead of 8 x 22) - the filter is shared-memory-load bound, not FMA bound.
  float gk[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) gk[k] = c_gauss[k];
  for (int task = threadIdx.x; task < kSI * (kST / 8); task += blockDim.x) {
    const int r = task / (kST / 8), c0 = (task - r * (kST / 8)) * 8;
    float pv[18], tv[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) { pv[k] = s_p[r][c0 + k]; tv[k] = s_t[r][c0 + k]; }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = gk[k], pp = pv[o + k], tt = tv[o + k];
        a0 += g * pp; a1 += g * tt; a2 += g * pp * pp; a3 += g * tt * tt; a4 += g * pp * tt;
      }
      s_h[0][r][c0 + o] = a0; s_h[1][r][c0 + o] = a1; s_h[2][r][c0 + o] = a2; s_h[3][r][c0 + o] = a3; s_h[4][r][c0 + o] = a4;
    }
  }
  __syncthreads();
  const float c1 = c12[0], c2 = c12[1];
  double local = 0;
  // Vertical pass: one thread = 4 consecutive output rows of one column (14 loads per map instead of 4 x 11).
  for (int task = threadIdx.x; task < kST * (kST / 4); task += blockDim.x) {
    const int c = task % kST, r0 = (task / kST) * 4;
    float m[4][5];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int q = 0; q < 5; ++q) m[o][q] = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float col[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) col[k] = s_h[q][r0 + k][c];
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k = 0; k < 11; ++k) m[o][q] += gk[k] * col[o + k];
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int y = oy + r0 + o, x = ox + c;
      if (y >= H - 5 || x >= W - 5) continue;
      const float mu_p2 = m[o][0] * m[o][0], mu_t2 = m[o][1] * m[o][1], mu_pt = m[o][0] * m[o][1];
      const float sp = m[o][2] - mu_p2, st = m[o][3] - mu_t2, spt = m[o][4] - mu_pt;
      local += static_cast<double>(((2.f * mu_pt + c1) * (2.f * spt + c2)) / ((mu_p2 + mu_t2 + c1) * (sp + st + c2)));
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += s_red[w];
    partials[(static_cast<long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = v;
  }
}

// Single block: SSIM partial sum + closed forms of the 16 metrics and the two losses.
__global__ void __launch_bounds__(256)
metrics_final_kernel(const double* __restrict__ totals, const double* __restrict__ ssim_partials, int n_ssim, double n_pix,
                     double n_ssim_pix, float* __restrict__ out) {
  __shared__ double sh[256];
  double part = 0;
  for (int i = threadIdx.x; i < n_ssim; i += 256) part += ssim_partials[i];      // fixed assignment + fixed tree: deterministic
  sh[threadIdx.x] = part;
  __syncthreads();
  for (int stride = 128; stride > 0; stride >>= 1) {
    if (threadIdx.x < stride) sh[threadIdx.x] += sh[threadIdx.x + stride];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const double ssim = sh[0];
  const double* t = totals;
  // RegressionAccuracy.compute: correct.float() / total - a float32 division of the rounded int64 counters
  for (int k = 0; k < 8; ++k) out[k] = __fdiv_rn(static_cast<float>(t[k]), static_cast<float>(n_pix));
  const double mse = t[9] / n_pix;
  const double tmin = fmin(t[kSums + 0], 0.0), tmax = fmax(t[kSums + 1], 0.0);       // PSNR states start at 0.0
  const double range = tmax - tmin;
  out[8] = static_cast<float>(10.0 * log10(range * range / mse));
  out[9] = static_cast<float>(ssim / n_ssim_pix);
  out[10] = static_cast<float>(t[8] / n_pix);
  out[11] = static_cast<float>(mse);
  out[12] = static_cast<float>(sqrt(mse));
  out[13] = static_cast<float>(t[13] / n_pix);
  out[14] = static_cast<float>(t[12] / n_pix);
  out[15] = static_cast<float>(1.0 - t[9] / (t[11] - t[10] * t[10] / n_pix));         // R2Score, flattened
  out[16] = static_cast<float>(t[14] / n_pix);                                        // nn.L1Loss (task.py:141)
  out[17] = static_cast<float>(t[15] / n_pix);                                        // nn.MSELoss
}

int p1_blocks_per_image(int n, long hw) {
  long want = (hw / 4 + kP1Threads * 4 - 1) / (kP1Threads * 4);   // >= 4 groups per thread
  long cap = (148L * 4 + n - 1) / n;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

}  // namespace

size_t metrics_scratch_bytes(int n, int h, int w) {
  const long hw = static_cast<long>(h) * w;
  const int bpi = p1_blocks_per_image(n, hw);
  const long ssim_blocks = static_cast<long>(n) * ((h - 10 + kST - 1) / kST) * ((w - 10 + kST - 1) / kST);
  return static_cast<size_t>(bpi) * n * kPart * 8 + kPart * 8 + 64 + (ssim_blocks > 0 ? ssim_blocks : 1) * 8 + 256;
}

cudaError_t launch_masked_metrics(const float* sr, const float* hr, const float* orig, const float* mask, const double* mn, const double* mx,
                                  float zmean, float zstd, double ra, double rb, double eps, int n, int h, int w, float* out, void* scratch,
                                  cudaStream_t s, int* launches) {
  // __constant__ memory is per device: upload the window once per device
  static bool gauss_ready[64] = {};
  int dev = 0;
  cudaError_t ed = cudaGetDevice(&dev);
  if (ed != cudaSuccess) return ed;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!gauss_ready[dev]) {
    // torchmetrics _gaussian: exp(-(d/sigma)^2/2) normalised, d = -5..5, sigma = 1.5
    float g[11];
    double sum = 0;
    for (int i = 0; i < 11; ++i) { g[i] = static_cast<float>(exp(-((i - 5) / 1.5) * ((i - 5) / 1.5) / 2.0)); sum += g[i]; }
    for (int i = 0; i < 11; ++i) g[i] = static_cast<float>(g[i] / sum);
    cudaError_t e = cudaMemcpyToSymbol(c_gauss, g, sizeof(g));
    if (e != cudaSuccess) return e;
    gauss_ready[dev] = true;
  }
  const long hw = static_cast<long>(h) * w;
  const int bpi = p1_blocks_per_image(n, hw);
  const int gy = (h - 10 + kST - 1) / kST, gx = (w - 10 + kST - 1) / kST;
  const int ssim_blocks = n * gy * gx;
  uint8_t* base = reinterpret_cast<uint8_t*>(scratch);
  double* partials = reinterpret_cast<double*>(base);
  double* totals = partials + static_cast<long>(bpi) * n * kPart;
  float* c12 = reinterpret_cast<float*>(totals + kPart);
  double* ssim_part = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(c12) + 64);
  if (mn) metrics_pass1_kernel<double><<<dim3(bpi, n), kP1Threads, 0, s>>>(sr, hr, orig, mask, mn, mx, zmean, zstd, ra, rb, eps, hw, partials);
  else metrics_pass1_kernel<float><<<dim3(bpi, n), kP1Threads, 0, s>>>(sr, hr, orig, mask, mn, mx, zmean, zstd, ra, rb, eps, hw, partials);
  metrics_reduce1_kernel<<<1, 256, 0, s>>>(partials, bpi * n, totals, c12);
  ssim_kernel<<<dim3(gx, gy, n), 256, 0, s>>>(sr, hr, mask, h, w, c12, ssim_part);
  metrics_final_kernel<<<1, 256, 0, s>>>(totals, ssim_part, ssim_blocks, static_cast<double>(hw) * n,
                                        static_cast<double>(n) * (h - 10) * (w - 10), out);
  *launches = 4;
  return cudaGetLastError();
}

}  // namespace csr
