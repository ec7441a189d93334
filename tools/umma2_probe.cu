// Bring-up probe for the CTA-pair MMA (tcgen05.mma.cta_group::2): which operand halves come from which CTA, where the
// accumulator rows land, and what an M=256 MMA costs.  Cluster of two CTAs; values are small integers (exact in bf16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_probe tools/umma2_probe.cu && ./umma2_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../climate-super-resolution_b200/csrc/ptx.cuh"
using namespace csr;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

// operand tile in shared memory: rows x 16 (K), K-major, no swizzle: [row group of 8][k chunk 2][8 rows][8 elems]
__device__ void fill(uint8_t* dst, int rows, int row0, int mul_r, int mul_k, int mod, int off) {
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  for (int i = threadIdx.x; i < rows * 16; i += blockDim.x) {
    const int e = i & 7, r8 = (i >> 3) & 7, kc = (i >> 6) & 1, g = i >> 7;
    const int r = row0 + g * 8 + r8, k = kc * 8 + e;
    d[i] = __float2bfloat16_rn(static_cast<float>((r * mul_r + k * mul_k) % mod - off));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(int N, int iters, float* out, long long* clk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t rank = cluster_rank();
  const int half = N / 2;
  fill(gen, 128, 128 * rank, 7, 3, 13, 6);                       // A rows of this CTA
  fill(gen + 8192, half, half * rank, 5, 1, 11, 5);               // B rows (N half) of this CTA
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc2(smem_u32(&tslot), 256); tmem_relinquish2(); }
  tc_fence_before(); __syncthreads(); cluster_sync(); tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t idesc = make_idesc_bf16(256, N);
  const uint32_t hi = (256u >> 4) | (1u << 14);                  // SBO 256, version 1, no swizzle
  const uint32_t a_lo = ((base >> 4) & 0x3FFF) | (8u << 16);     // LBO 128
  const uint32_t b_lo = (((base + 8192) >> 4) & 0x3FFF) | (8u << 16);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    if (rank == 0 && elect_one()) { umma2(tmem, a_lo, hi, b_lo, hi, idesc, 0); commit2(smem_u32(&bar), 3); }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  // read back this CTA's 128 lanes x N columns
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < N; c += 8) {
      uint32_t r[8];
      tmem_ld8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out[(static_cast<size_t>(rank) * 128 + warp * 32 + lane) * 256 + c + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before(); __syncthreads(); cluster_sync(); tc_fence_after();
  if (iters > 0 && threadIdx.x < 32) {
    uint32_t phase = 1;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (rank == 0 && elect_one()) {
#pragma unroll
        for (int i = 0; i < 24; ++i) umma2(tmem, a_lo, hi, b_lo, hi, idesc, 1);
        commit2(smem_u32(&bar), 3);
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase); phase ^= 1;
    }
    t1 = clock64();
  }
  tc_fence_before(); __syncthreads(); cluster_sync();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc2(tmem, 256); }
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* d; long long* c;
  cudaMalloc(&d, 2 * 128 * 256 * 4 * 74); cudaMalloc(&c, 148 * 8);
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {64, 192}) {
    cudaMemset(d, 0, 2 * 128 * 256 * 4);
    probe2<<<2, 128, 64 * 1024>>>(N, 0, d, c);
    cudaError_t e = cudaDeviceSynchronize();
    printf("N=%d correctness launch: %s\n", N, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> h(2 * 128 * 256);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    auto A = [](int r, int k) { return (float)((r * 7 + k * 3) % 13 - 6); };
    auto B = [](int n, int k) { return (float)((n * 5 + k) % 11 - 5); };
    // hypothesis H0: out[rank][lane][col] = sum_k A[128*rank+lane][k] * B[col][k], B rows [0,N/2) from CTA0, [N/2,N) from CTA1
    long bad = 0; int first = -1;
    for (int rk = 0; rk < 2; ++rk) for (int l = 0; l < 128; ++l) for (int col = 0; col < N; ++col) {
      float ref = 0; for (int k = 0; k < 16; ++k) ref += A(128 * rk + l, k) * B(col, k);
      const float got = h[(rk * 128 + l) * 256 + col];
      if (got != ref) { if (first < 0) first = (rk * 128 + l) * 256 + col; ++bad; }
    }
    printf("N=%d H0 (rows by CTA, B N-halves by CTA): %ld mismatches of %d", N, bad, 256 * N);
    if (first >= 0) printf(" first at rank %d lane %d col %d got %g", first / (128 * 256), (first / 256) % 128, first % 256, h[first]);
    printf("\n");
    if (bad) {
      // dump a few values for manual analysis
      for (int rk = 0; rk < 2; ++rk) { printf(" rank %d lane 0 cols: ", rk); for (int col = 0; col < 16; ++col) printf("%g ", h[(rk * 128) * 256 + col]); printf("\n"); }
      for (int col : {0, 1, N / 2, N / 2 + 1}) {
        float r0 = 0, r128 = 0; for (int k = 0; k < 16; ++k) { r0 += A(0, k) * B(col, k); r128 += A(128, k) * B(col, k); }
        printf(" ref row0 col%d %g ; row128 col%d %g\n", col, r0, col, r128);
      }
    }
  }
  // timing: 74 clusters, N=192, M=256
  for (int N : {192, 128, 256}) {
    probe2<<<148, 128, 64 * 1024>>>(N, 200, d, c);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, c, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; i += 2) mx = h[i] > mx ? h[i] : mx;
    printf("timing M=256 N=%d: %.1f clk per MMA (per-SM tensor floor %.1f) %s\n", N, (double)mx / (200 * 24.0), N / 2.0, cudaGetErrorString(e));
  }
  return 0;
}
