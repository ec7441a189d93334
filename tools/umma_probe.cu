// Microbenchmark: cost per tcgen05.mma (M=128, K=16, bf16, SS mode) as a function of N, B layout and how the
// descriptors are produced.  One CTA per SM, operands are zeros in shared memory (values do not matter).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "../climate-super-resolution_b200/csrc/ptx.cuh"
using namespace csr;

template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, int b_swizzled, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); fence_proxy_async_smem(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t idesc = make_idesc_bf16(128, N);
  const uint32_t a_addr = base, b_addr = base + 64 * 1024;
  const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t b_hi = b_swizzled ? a_hi : ((256u >> 4) | (1u << 14));
  const uint32_t a_lo0 = (a_addr >> 4) | (1u << 16);
  const uint32_t b_lo0 = (b_addr >> 4) | ((b_swizzled ? 1u : 8u) << 16);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    uint32_t phase = 0;
    // warm-up
    if (elect_one()) { umma_bf16_split(tmem, a_lo0, a_hi, b_lo0, b_hi, idesc, 0); umma_commit(smem_u32(&bar)); }
    __syncwarp(); mbar_wait(smem_u32(&bar), phase); phase ^= 1;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        if (MODE == 0) {          // 36 MMAs, constant descriptors
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_split(tmem, a_lo0, a_hi, b_lo0, b_hi, idesc, 1);
        } else if (MODE == 1) {   // 36 MMAs, conv-like address pattern (3x3 taps x 4 k-steps), fully unrolled
          uint32_t b_lo = b_lo0;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) { umma_bf16_split(tmem, a_lo0 + (dy * 16 + dx) * 8 + ks * 2, a_hi, b_lo, b_hi, idesc, 1); b_lo += (N * 32) >> 4; }
        } else {                  // 36 MMAs, alternate two accumulators (independent destinations)
#pragma unroll
          for (int i = 0; i < 36; ++i) umma_bf16_split(tmem + (i & 1) * 256, a_lo0, a_hi, b_lo0, b_hi, idesc, 1);
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase); phase ^= 1;
    }
    t1 = clock64();
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int bsw = 0; bsw < 2; ++bsw)
    for (int N : {16, 32, 64, 128, 256}) {
      const int iters = 200;
      probe<MODE><<<grid, 128, 200 * 1024>>>(N, iters, bsw, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-28s grid %3d B=%s N=%3d: %7.1f clk/MMA (ideal tensor %5.1f)  %s\n", name, grid, bsw ? "sw128" : "nosw ", N,
             (double)mx / (iters * 36.0), N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<0>("const-desc", grid);
    run<1>("conv-pattern", grid);
    run<2>("two-accumulators", grid);
  }
  return 0;
}
