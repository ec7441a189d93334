// KHxKW stride-1 convolution over NHWC bf16 as an implicit GEMM on tcgen05 (sm_100a).
//
// Measured on B200 (profiles/r01_umma_probe.txt): in SS mode an M=128,K=16 bf16 MMA costs ~45-55 clk for any N <= 64
// (the 4 KB A-operand read from shared memory is the floor), ~71 clk at N=128, ~135 at N=256.  A conv whose GEMM N is
// only Cout = 16..64 therefore wastes most of each MMA.  This kernel folds the KW horizontal taps into N:
//
//   D[128 window pixels][dx*npad + co] (fp32, TMEM) += A_dy[128 pixels][16 ch] x B_dy[KW*npad][16 ch]
//
// for every vertical tap dy and 16-channel k-step: KH (not KH*KW) MMAs per k-step, each KW times wider.  The epilogue
// finishes the horizontal sum with warp shuffles:  out[m] = sum_dx D[m + dx - PW][dx*npad + co]  (GEMM rows m are
// flattened window positions, one TMEM lane == one thread per row, neighbours in x are neighbour lanes; SW <= 32 so a
// window row never straddles two warps).
//
// Halo reuse: one TMA box per 64-channel k-block brings the (TH+KH-1) x SW pixel window of a tile into shared memory
// ONCE (out-of-bounds pixels are zero-filled by TMA == the conv's zero padding).  The A operand of vertical tap dy is
// the same window read from byte offset dy*SW*128: only the descriptor start address changes between taps (UMMA
// applies the 128B swizzle on absolute smem address bits, so any 128-byte-aligned start works - verified on HW).
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner, warps 2..17 = epilogue (four warps
// per TMEM lane quadrant, each owning every 4th 8-channel chunk).  Accumulators are double buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1; the layer's packed weights stay resident in shared memory for the
// whole persistent CTA.  bf16 outputs are staged in a (swizzled) shared-memory tile and written with one bulk tensor
// store per tile (TMA clips rows/columns outside the image; a strided output map scatters a sub-pixel phase of the
// nearest-x2 + conv layers).  Programmatic dependent launch: everything before griddepcontrol.wait (barrier init, TMEM
// alloc, weight/bias loads) overlaps the tail of the previous layer's kernel.
//
// Replaces one nn.Conv2d(+LeakyReLU/ReLU, *0.2+x, cat, nearest-x2) call site of the reference generator:
// climsr/models/esrgan.py:33-38, 50-54, 90-100 and climsr/models/srcnn.py:14-16 - and, with transposed / flipped
// weight packs, the input-gradient half of their autograd backward.
#include <cstdio>

#include "conv_tc.cuh"
#include "ptx.cuh"

namespace csr {

namespace {

#define CSR_TRACE(role, tile_it, ev)                                                                                  \
  do {                                                                                                                \
    if (p.trace && blockIdx.x == 0 && (tile_it) < 64) p.trace[((role) * 64 + (tile_it)) * 4 + (ev)] = clock64();      \
  } while (0)

struct Tile {
  int n, y0, x0;
};

__device__ __forceinline__ Tile decode_tile(const ConvParams& p, int t) {
  const int per_img = p.tiles_x * p.tiles_y;
  Tile r;
  r.n = t / per_img;
  const int rem = t - r.n * per_img;
  const int ty = rem / p.tiles_x;
  r.y0 = ty * p.TH;
  r.x0 = (rem - ty * p.tiles_x) * p.TW;
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return v >= 0.f ? v : 0.2f * v;
  if (act == 2) return fmaxf(v, 0.f);
  return v;
}

// acc[j] += value of raw[j] held by lane (lane + delta); delta == 0 -> own value.  Executed by all 32 lanes.
__device__ __forceinline__ void gather_add8(float (&acc)[8], const uint32_t (&raw)[8], int delta, int lane) {
  if (delta == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __uint_as_float(raw[j]);
  } else {
    const int src = (lane + delta) & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __uint_as_float(__shfl_sync(0xffffffffu, raw[j], src));
  }
}

__device__ __forceinline__ void fma_residual8(float (&v)[8], const uint4& r, float scale) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = v[2 * i] * scale + bf16lo(w[i]);
    v[2 * i + 1] = v[2 * i + 1] * scale + bf16hi(w[i]);
  }
}

__device__ __forceinline__ void gate8(float (&v)[8], const uint4& g, float neg) {
  const uint32_t w[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] *= (bf16lo(w[i]) > 0.f) ? 1.f : neg;
    v[2 * i + 1] *= (bf16hi(w[i]) > 0.f) ? 1.f : neg;
  }
}

__device__ __forceinline__ uint4 ldg16(const void* base, size_t pix, int C, int coff) {
  return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + pix * C + coff));
}

}  // namespace

template <int KW_T>   // compile-time horizontal tap count (0 = runtime p.KW, taps gathered one at a time)
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: the 128B-swizzle pattern repeats every 8 rows x 128 B.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slots_addr = smem_base;
  const uint32_t stage_addr = slots_addr + static_cast<uint32_t>(p.n_slots) * p.slot_bytes;   // 2 staging buffers (TMA store)
  const uint32_t w_addr = stage_addr + 2u * p.stage_bytes;
  const uint32_t bias_addr = w_addr + ((p.w_bytes + 127) & ~127);
  const uint32_t bar_addr = bias_addr + 256;            // up to 64 fp32 biases
  // barriers: [0] weights, [1..S] a_full, [1+S..2S] a_empty, then acc_full[2], acc_empty[2]
  const int S = p.n_slots;
  auto bar_w = bar_addr;
  auto bar_a_full = [&](int s) { return bar_addr + 8u * (1 + s); };
  auto bar_a_empty = [&](int s) { return bar_addr + 8u * (1 + S + s); };
  auto bar_acc_full = [&](int b) { return bar_addr + 8u * (1 + 2 * S + b); };
  auto bar_acc_empty = [&](int b) { return bar_addr + 8u * (3 + 2 * S + b); };
  const uint32_t tmem_slot_addr = bar_addr + 8u * (5 + 2 * S);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (bias_addr - smem_base));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot_addr - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KW = KW_T ? KW_T : p.KW;

  // Let the next layer's grid start its own prologue as early as the hardware allows (PDL).
  griddep_launch_dependents();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    tma_prefetch_desc(&tmap_out);
    mbar_init(bar_w, 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_a_full(s), 1);
      mbar_init(bar_a_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), kEpilogueWarps);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
    // layer weights (constant data, not produced by the previous kernel): resident for the whole CTA
    mbar_arrive_expect_tx(bar_w, p.w_bytes);
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpk);
    for (int off = 0; off < p.w_bytes; off += 32768) {
      const int nbytes = min(32768, p.w_bytes - off);
      bulk_load(w_addr + off, wsrc + off, nbytes, bar_w);
    }
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot_addr, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.npad; i += blockDim.x) bias_s[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Everything below reads / writes activations of the previous layer(s): wait for the prerequisite grid.
  griddep_wait();

  const int ksteps_total = p.cin >> 4;
  const int nmma = KW * p.npad;                          // UMMA N: horizontal taps folded into the output columns

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    int slot = 0, pit = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++pit) {
      const Tile tl = decode_tile(p, t);
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        if (kb == 0 && lane == 0) CSR_TRACE(0, pit, 0);
        mbar_wait(bar_a_empty(slot), phase ^ 1);
        if (kb == 0 && lane == 0) CSR_TRACE(0, pit, 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_a_full(slot), p.win_bytes);
          tma_load_4d(slots_addr + slot * p.slot_bytes, &tmap, bar_a_full(slot), p.cin_off + kb * 64, tl.x0 - p.PW, tl.y0 - p.PH,
                      tl.n);
        }
        __syncwarp();
        if (++slot == S) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the warp-uniform control flow; one elected lane issues.  Per MMA only the 14-bit start-address
    // fields of the two descriptors change.
    const uint32_t idesc = make_idesc_bf16(kTileM, nmma);
    const uint32_t b_step16 = static_cast<uint32_t>(nmma * 32) >> 4;          // one (dy,kstep) weight block in 16-byte units
    const uint32_t kb_w16 = static_cast<uint32_t>(p.KH * 4) * b_step16;       // one full k-block of weights
    const uint32_t row16 = static_cast<uint32_t>(p.SW) * 8u;                  // one window row (SW pixels x 128 B)
    const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);             // SBO 1024 B, version 1, 128B swizzle
    const uint32_t b_hi = (256u >> 4) | (1u << 14);                           // SBO 256 B, version 1, no swizzle
    const uint32_t a_lbo = (16u >> 4) << 16, b_lbo = (128u >> 4) << 16;
    mbar_wait(bar_w, 0);
    tc_fence_after();
    int slot = 0, it = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      if (lane == 0) CSR_TRACE(1, it, 0);
      mbar_wait(bar_acc_empty(buf), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) CSR_TRACE(1, it, 1);
      const uint32_t d_tmem = tmem_base + buf * nmma;
      for (int kb = 0; kb < p.n_kblocks; ++kb) {
        mbar_wait(bar_a_full(slot), phase);
        tc_fence_after();
        if (kb == 0 && lane == 0) CSR_TRACE(1, it, 2);
        const int ks_here = min(4, ksteps_total - kb * 4);
        uint32_t a16 = ((slots_addr + slot * p.slot_bytes) >> 4) | a_lbo;
        uint32_t b16 = ((w_addr >> 4) + static_cast<uint32_t>(kb) * kb_w16) | b_lbo;
        if (elect_one()) {
          uint32_t acc = kb ? 1u : 0u;
          for (int dy = 0; dy < p.KH; ++dy, a16 += row16) {
            for (int ks = 0; ks < ks_here; ++ks, b16 += b_step16) {
              umma_bf16_split(d_tmem, a16 + ks * 2, a_hi, b16, b_hi, idesc, acc);
              acc = 1;
            }
          }
          umma_commit(bar_a_empty(slot));                                  // window slot reusable once these MMAs have read it
          if (kb == p.n_kblocks - 1) umma_commit(bar_acc_full(buf));       // accumulator complete -> epilogue
        }
        __syncwarp();
        if (kb == p.n_kblocks - 1 && lane == 0) CSR_TRACE(1, it, 3);
        if (++slot == S) { slot = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..17; TMEM lane quadrant = warp % 4, chunk group = (warp-2) / 4 ==========
    const int lane_grp = warp & 3;
    const int grp = (warp - 2) >> 2;                     // 0..3: owns 8-channel chunks grp, grp+4
    const bool issuer = (threadIdx.x == 64);             // first epilogue thread issues the bulk tensor stores
    const int m = lane_grp * 32 + lane;
    const int ty = m >> p.sw_shift;
    const int tx = m & (p.SW - 1);                       // window column
    const int n_chunks = p.npad >> 3;
    const bool col_ok = (tx >= p.PW) && (tx < p.PW + p.TW);
    const int srow = ty * p.TW + (tx - p.PW);            // staging row of this thread's pixel
    const bool tma_store = (p.store_mode == kStoreTma);
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const Tile tl = decode_tile(p, t);
      const int buf = it & 1;
      const int y = tl.y0 + ty;
      const int x = tl.x0 - p.PW + tx;
      const bool valid = col_ok && (y < p.H) && (x < p.W);
      const size_t pix = (static_cast<size_t>(tl.n) * p.H + y) * p.W + x;
      // residual / gate operands do not depend on the accumulator: fetch them before waiting for the MMAs
      uint4 q1[2], q2[2], qg[2];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int ch0 = (grp + 4 * ci) * 8;
        const bool on = valid && (ch0 < p.n_store);
        q1[ci] = (on && p.r1) ? ldg16(p.r1, pix, p.r1_C, p.r1_coff + ch0) : make_uint4(0, 0, 0, 0);
        q2[ci] = (on && p.r2) ? ldg16(p.r2, pix, p.r2_C, p.r2_coff + ch0) : make_uint4(0, 0, 0, 0);
        qg[ci] = (on && p.gate && ch0 >= p.gate_from) ? ldg16(p.gate, pix, p.gate_C, p.gate_coff + ch0)
                                                       : make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
      }
      if (threadIdx.x == 64) CSR_TRACE(2, it, 0);
      mbar_wait(bar_acc_full(buf), (it >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) CSR_TRACE(2, it, 1);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + buf * nmma;
      float v[2][8];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = grp + 4 * ci;
        if (c < n_chunks) {                               // warp-uniform
          const int ch0 = c * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) v[ci][i] = bias_s[ch0 + i];
          if constexpr (KW_T == 1) {
            uint32_t r0[8];
            tmem_ld8(t_addr + ch0, r0);
            tmem_ld_wait();
            gather_add8(v[ci], r0, 0, lane);
          } else if constexpr (KW_T == 2) {
            uint32_t r0[8], r1[8];
            tmem_ld8(t_addr + ch0, r0);
            tmem_ld8(t_addr + p.npad + ch0, r1);
            tmem_ld_wait();
            gather_add8(v[ci], r0, -p.PW, lane);
            gather_add8(v[ci], r1, 1 - p.PW, lane);
          } else if constexpr (KW_T == 3) {
            uint32_t r0[8], r1[8], r2[8];
            tmem_ld8(t_addr + ch0, r0);
            tmem_ld8(t_addr + p.npad + ch0, r1);
            tmem_ld8(t_addr + 2 * p.npad + ch0, r2);
            tmem_ld_wait();
            gather_add8(v[ci], r0, -p.PW, lane);
            gather_add8(v[ci], r1, 1 - p.PW, lane);
            gather_add8(v[ci], r2, 2 - p.PW, lane);
          } else {
            for (int dx = 0; dx < KW; ++dx) {
              uint32_t r[8];
              tmem_ld8(t_addr + dx * p.npad + ch0, r);
              tmem_ld_wait();
              gather_add8(v[ci], r, dx - p.PW, lane);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[ci][i] = apply_act(v[ci][i], p.act);
          if (p.r1) fma_residual8(v[ci], q1[ci], p.s1);
          if (p.r2) fma_residual8(v[ci], q2[ci], p.s2);
          if (p.gate) gate8(v[ci], qg[ci], p.gate_neg);
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty(buf));

      if (tma_store) {
        const uint32_t sbuf = stage_addr + static_cast<uint32_t>(it & 1) * p.stage_bytes;
        if (issuer) bulk_wait_group_read<1>();            // the store issued two tiles ago has finished reading this buffer
        named_bar_sync(1, kEpilogueThreads);
        if (col_ok) {
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            const int c = grp + 4 * ci;
            if (c * 8 < p.n_store) {
              const int cs = (p.stage_row_bytes == 128) ? (c ^ (srow & 7)) : c;    // SWIZZLE_128B staging for 64-channel rows
              st_shared_v4(sbuf + srow * p.stage_row_bytes + cs * 16, pack_bf16x2(v[ci][0], v[ci][1]), pack_bf16x2(v[ci][2], v[ci][3]),
                           pack_bf16x2(v[ci][4], v[ci][5]), pack_bf16x2(v[ci][6], v[ci][7]));
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, kEpilogueThreads);
        if (issuer) {
          tma_store_4d(&tmap_out, sbuf, p.out_coff, tl.x0, tl.y0, tl.n);
          bulk_commit_group();
        }
      } else if (valid) {
        const size_t opix = (static_cast<size_t>(tl.n) * p.out_H + (y * p.out_sy + p.out_oy)) * p.out_W + (x * p.out_sx + p.out_ox);
        if (p.store_mode == kStoreF32Planar) {
          if (grp == 0) reinterpret_cast<float*>(p.out)[opix] = v[0][0];
        } else {
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            const int ch0 = (grp + 4 * ci) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (ch0 + i < p.n_store) {
                if (p.store_mode == kStoreF32Nhwc)
                  reinterpret_cast<float*>(p.out)[opix * p.out_C + p.out_coff + ch0 + i] = v[ci][i];
                else
                  reinterpret_cast<__nv_bfloat16*>(p.out)[opix * p.out_C + p.out_coff + ch0 + i] = __float2bfloat16_rn(v[ci][i]);
              }
            }
          }
        }
      }
      if (threadIdx.x == 64) CSR_TRACE(2, it, 2);
    }
    if (tma_store && issuer) bulk_wait_group<0>();        // all bulk stores of this CTA have completed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

size_t conv_smem_bytes(const ConvParams& p) {
  return 1024 /*alignment slack*/ + static_cast<size_t>(p.n_slots) * p.slot_bytes + 2 * static_cast<size_t>(p.stage_bytes) +
         ((p.w_bytes + 127) & ~127) + 256 /*bias*/ + 8 * (5 + 2 * p.n_slots) + 16;
}

template <int KW_T>
static int launch_t(const ConvParams& p, const CUtensorMap& tmap, const CUtensorMap& tmap_out, int num_sms, cudaStream_t stream) {
  const size_t smem = conv_smem_bytes(p);
  if (smem > static_cast<size_t>(kSmemLimit)) return static_cast<int>(cudaErrorInvalidValue);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<KW_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.num_tiles < num_sms ? p.num_tiles : num_sms);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = p.use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, conv_tc_kernel<KW_T>, p, tmap, tmap_out));
}

int launch_conv_tc(const ConvParams& p, const CUtensorMap& tmap, const CUtensorMap& tmap_out, int num_sms, cudaStream_t stream) {
  switch (p.KW) {
    case 1: return launch_t<1>(p, tmap, tmap_out, num_sms, stream);
    case 2: return launch_t<2>(p, tmap, tmap_out, num_sms, stream);
    case 3: return launch_t<3>(p, tmap, tmap_out, num_sms, stream);
    default: return launch_t<0>(p, tmap, tmap_out, num_sms, stream);
  }
}

}  // namespace csr
