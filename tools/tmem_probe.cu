// Microbenchmarks behind the round-2 epilogue redesign (results: profiles/r02_tmem_probe.txt):
//   1. tcgen05.ld throughput per SM as a function of the number of reading warps, the .xN width and how many loads are
//      in flight before tcgen05.wait::ld - is a 128 x 192-column accumulator tile bound by the TMEM read port?
//   2. the same while another warp keeps the tensor pipe busy with N = 192 MMAs
//   3. MMA cost of the tap-folding variants of a 3x3 conv with 64 output channels, per (dy, k-step):
//        one N=192 MMA (all three horizontal taps folded into N)            -> 192 accumulator columns to read back
//        one N=128 MMA + one N=64 MMA whose A start is shifted by one pixel   -> 128 columns
//        three N=64 MMAs with A shifted by 0/1/2 pixels (no folding)          -> 64 columns, no shuffles in the epilogue
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tools/tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "../climate-super-resolution_b200/csrc/ptx.cuh"
using namespace csr;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// X = columns per load (8/16/32), DEPTH = loads issued before one wait::ld.  MMA_BG: warp `nwarps` issues N=192 MMAs meanwhile.
template <int X, int DEPTH>
__global__ void __launch_bounds__(544, 1) ldprobe(int nwarps, int iters, int mma_bg, long long* out, unsigned* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { stop = 0; mbar_init(smem_u32(&bar), 1); fence_mbar_init(); fence_proxy_async_smem(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const int warp = threadIdx.x >> 5;
  long long t0 = 0, t1 = 0;
  unsigned acc = 0;
  if (warp < nwarps) {
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c = 0; c < 512; c += X * DEPTH) {
        if constexpr (X == 8) {
          uint32_t r[DEPTH][8];
#pragma unroll
          for (int d = 0; d < DEPTH; ++d) tmem_ld8(t_lane + c + d * 8, r[d]);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc ^= r[d][j];
        } else if constexpr (X == 16) {
          uint32_t r[DEPTH][16];
#pragma unroll
          for (int d = 0; d < DEPTH; ++d) tmem_ld16(t_lane + c + d * 16, r[d]);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= r[d][j];
        } else {
          uint32_t r[DEPTH][32];
#pragma unroll
          for (int d = 0; d < DEPTH; ++d) tmem_ld32(t_lane + c + d * 32, r[d]);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[d][j];
        }
      }
    }
    t1 = clock64();
    if (warp == 0 && (threadIdx.x & 31) == 0) stop = 1;
  } else if (warp == 16 && mma_bg) {
    const uint32_t idesc = make_idesc_bf16(128, 192);
    const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = (256u >> 4) | (1u << 14);
    const uint32_t a_lo0 = (base >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((base + 32 * 1024) >> 4) | (8u << 16);
    uint32_t phase = 0;
    long long n = 0;
    t0 = clock64();
    while (!stop) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 12; ++i) umma_bf16_split(tmem + (i & 1) * 256, a_lo0, a_hi, b_lo0, b_hi, idesc, 1);
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase); phase ^= 1;
      n += 12;
    }
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[160 + blockIdx.x] = (t1 - t0) / (n ? n : 1);
  }
  if (sink && acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int X, int DEPTH>
void run_ld(const char* name) {
  long long* d; cudaMalloc(&d, 512 * 8);
  cudaFuncSetAttribute(ldprobe<X, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int bg = 0; bg < 2; ++bg)
    for (int nw : {1, 4, 8, 16}) {
      const int iters = 64;
      cudaMemset(d, 0, 512 * 8);
      ldprobe<X, DEPTH><<<148, 544, 100 * 1024>>>(nw, iters, bg, d, nullptr);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[512]; cudaMemcpy(h, d, 512 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)nw * 32 * 512 * 4 * iters;
      printf("%-18s warps %2d mma_bg %d: %7.1f B/clk/SM  (128x192 fp32 tile = %6.0f clk)", name, nw, bg, bytes / mx, 128.0 * 192 * 4 / (bytes / mx));
      if (bg) printf("   bg N=192 MMA: %lld clk/MMA", h[160]);
      printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  cudaFree(d);
}

// MODE 0: 12 x N=192; 1: 12 x (N=128 + N=64 shifted A); 2: 36 x N=64 (A shifted 0/1/2 px); 3: 12 x N=48 x3 (thin, folded) ; 4: 12 x N=144
template <int MODE>
__global__ void __launch_bounds__(128, 1) mmaprobe(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); fence_proxy_async_smem(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t b_hi = (256u >> 4) | (1u << 14);
  const uint32_t a_lo0 = (base >> 4) | (1u << 16);
  const uint32_t b_lo0 = ((base + 48 * 1024) >> 4) | (8u << 16);
  const uint32_t i192 = make_idesc_bf16(128, 192), i128 = make_idesc_bf16(128, 128), i64 = make_idesc_bf16(128, 64), i48 = make_idesc_bf16(128, 48),
                 i144 = make_idesc_bf16(128, 144);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    uint32_t phase = 0;
    if (elect_one()) { umma_bf16_split(tmem, a_lo0, a_hi, b_lo0, b_hi, i64, 0); umma_commit(smem_u32(&bar)); }
    __syncwarp(); mbar_wait(smem_u32(&bar), phase); phase ^= 1;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        uint32_t b = b_lo0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t a = a_lo0 + dy * 16 * 8 + ks * 2;   // window row dy (SW = 16 pixels of 128 B), k-step ks (32 B)
            if (MODE == 0) { umma_bf16_split(tmem, a, a_hi, b, b_hi, i192, 1); b += (192 * 32) >> 4; }
            else if (MODE == 1) {
              umma_bf16_split(tmem, a, a_hi, b, b_hi, i128, 1); b += (128 * 32) >> 4;
              umma_bf16_split(tmem + 64, a + 8, a_hi, b, b_hi, i64, 1); b += (64 * 32) >> 4;
            } else if (MODE == 2) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) { umma_bf16_split(tmem, a + dx * 8, a_hi, b, b_hi, i64, 1); b += (64 * 32) >> 4; }
            } else if (MODE == 3) {
              umma_bf16_split(tmem, a, a_hi, b, b_hi, i48, 1); b += (48 * 32) >> 4;
            } else {
              umma_bf16_split(tmem, a, a_hi, b, b_hi, i144, 1); b += (144 * 32) >> 4;
            }
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
void run_mma(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(mmaprobe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 200;
  mmaprobe<MODE><<<148, 128, 200 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-44s %7.1f clk per (dy, k-step)   %7.0f clk per 128 px x 64 ch-in tile  %s\n", name, (double)mx / (iters * 12.0), (double)mx / iters,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

// 4. does epilogue traffic slow the MMAs down?  Warp 16 issues N = NMMA MMAs back to back while `nwarps` other warps loop over
//    FG = 0 nothing, 1 warp shuffles, 2 16-byte shared-memory stores + loads (conflict-free), 3 FFMA only.
template <int NMMA, int FG>
__global__ void __launch_bounds__(544, 1) contend(int nwarps, int iters, long long* out, unsigned* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); fence_proxy_async_smem(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = 0, t1 = 0;
  if (warp == 16) {
    const uint32_t idesc = make_idesc_bf16(128, NMMA);
    const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = (256u >> 4) | (1u << 14);
    const uint32_t a_lo0 = (base >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((base + 48 * 1024) >> 4) | (8u << 16);
    uint32_t phase = 0;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 12; ++i) umma_bf16_split(tmem + (i & 1) * 256, a_lo0 + (i % 3) * 16 * 8 + (i & 3) * 2, a_hi, b_lo0 + i * ((NMMA * 32) >> 4), b_hi, idesc, 1);
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase); phase ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[blockIdx.x] = (t1 - t0) / (12LL * iters);
  } else if (warp < nwarps) {
    // ~ the same wall time as the MMA loop: iters * 12 * ~80 clk
    const uint32_t my = base + 64 * 1024 + threadIdx.x * 16;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * 0.5f + j;
    for (int it = 0; it < iters * 6; ++it) {
      if (FG == 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += __shfl_sync(0xffffffffu, v[(j + 1) & 7], (lane + 1) & 31);
      } else if (FG == 2) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          st_shared_v4(my, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
          const uint4 q = ld_shared_v4(my ^ 16u);
          v[0] += __uint_as_float(q.x); v[1] += __uint_as_float(q.y); v[2] += __uint_as_float(q.z); v[3] += __uint_as_float(q.w);
        }
      } else if (FG == 3) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], 1.0001f, v[(j + 1) & 7]);
      }
    }
    if (sink && v[0] + v[1] + v[2] + v[3] == 12345.f) sink[0] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int NMMA, int FG>
void run_contend(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(contend<NMMA, FG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int nw : {0, 4, 8, 16}) {
    contend<NMMA, FG><<<148, 544, 200 * 1024>>>(nw, 200, d, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%3d MMAs while %2d warps run %-28s %5lld clk per MMA  %s\n", NMMA, nw, name, mx, e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (FG == 0) break;
  }
  cudaFree(d);
}

int main() {
  run_contend<144, 0>("nothing");
  run_contend<144, 1>("32 shuffles per iteration");
  run_contend<144, 2>("4 x (STS.128 + LDS.128)");
  run_contend<144, 3>("32 FFMA per iteration");
  run_contend<48, 0>("nothing");
  run_contend<48, 1>("32 shuffles per iteration");
  run_contend<48, 2>("4 x (STS.128 + LDS.128)");
  run_mma<0>("3 taps folded: N=192");
  run_mma<1>("2 folded + 1 shifted: N=128 + N=64");
  run_mma<2>("unfolded: 3 x N=64, A shifted by 0/1/2 px");
  run_mma<3>("thin folded: N=48");
  run_mma<4>("N=144");
  run_ld<8, 1>("ld.x8 depth1");
  run_ld<8, 4>("ld.x8 depth4");
  run_ld<16, 1>("ld.x16 depth1");
  run_ld<16, 2>("ld.x16 depth2");
  run_ld<32, 1>("ld.x32 depth1");
  run_ld<32, 2>("ld.x32 depth2");
  return 0;
}
