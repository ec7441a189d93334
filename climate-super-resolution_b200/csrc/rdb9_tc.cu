// conv1..conv4 of a ResidualDenseBlock (climsr/models/esrgan.py:33-36, gc = 16) as ONE persistent dataflow launch with ALL NINE
// taps folded into the UMMA N dimension (N = 9 * 16 = 144).
//
// Why: the thin layers are bound by the operand fetch of their MMAs, not by tensor math (an M = 128, K = 16 MMA costs ~64 clk for any
// N <= 64, 93 clk at N = 144; profiles/r02_tmem_mma_probe.txt).  dense_block_kernel (rdb_tc.cu) folds the three horizontal taps into N
// (N = 48) and issues one MMA per vertical tap and k-step: 3 x 70 clk per k-step and tile.  Here one MMA per k-step computes
//     D[m][(dy, dx, co)] = sum_c in[m][c] * W[co][c][dy][dx]
// for every window position m, and the epilogue finishes  out(r, c) = sum_{dy,dx} D[(r + dy - 1, c + dx - 1)][(dy, dx, .)] :
// horizontal neighbours by warp shuffle (a 16-pixel window row is half a warp), the row above / below by one more shuffle (other
// half-warp) or through shared memory (neighbouring warp of the 4-warp epilogue group; neighbouring M tile of the window).
// Windows are 16 x 16 input pixels = two M tiles of 8 rows, 14 x 14 outputs (1-pixel halo all round), so fewer of the 128 MMA rows
// are useful than in the N = 48 kernel (50 instead of 40 tiles per 64 x 64 image), but a tile takes ~100 instead of 210 clk per k-step.
//
// The k-step blocks of the packed weights are those of conv_tc / rdb_tc ([k-block][dy][k-step] blocks of 48 rows x 16 K = 1536 bytes);
// the weight loader copies them to shared memory in [k-block][k-step][dy] order, which makes the 144 rows of a k-step contiguous.
// Dataflow protocol (per-(layer, window) counters, acquire / release, proxy fence once per window), window -> CTA map, weight
// double-buffering and deadlock argument are those of rdb_tc.cu.  The summation order differs from the per-layer kernels (nine taps
// summed in fp32 in the epilogue instead of three inside the accumulator), so results agree to fp32 rounding, not bit for bit.
//
// MEASURED AND REJECTED (round 2, profiles/r02_dense_block_notes.txt): correct (tests/test_gpu_parity.py, option 32), but 110 us per dense
// block at cfg2 against 52.7 us for rdb_tc.cu.  (a) TMEM holds only three 144-column accumulators and a window's two tiles are drained
// by ONE group (they share the row 7 / 8 boundary), so MMAs and epilogues of consecutive windows hardly overlap: ~1800 clk per tile =
// MMA + epilogue in series.  (b) With the epilogue stubbed out the MMA + TMA pipeline alone takes 40 us, not the 25 us of its MMAs: a
// dense block moves ~2.4 MB of window boxes per CTA from L2 to shared memory (every layer re-reads the block input, 64-channel boxes
// even for the 16..48 channels of the second k-block) = 9 TB/s over 148 SMs - the L2 roof, which also sits 37 us under rdb_tc.cu.
// Compiled only with CSR_EXPERIMENTS=1 (build.py).
#ifdef CSR_EXPERIMENTS
#include <cstdio>

#include "conv_tc.cuh"
#include "ptx.cuh"
#include "rdb_tc.cuh"

namespace csr {

namespace {

constexpr int kNpad = 16;             // output channels per layer (gc)
constexpr int kN9 = 9 * kNpad;        // UMMA N
constexpr int kAcc9 = 3;              // accumulator buffers (144 of 160 columns each)
constexpr int kAccStride = 160;
constexpr int kGroups = 4;            // epilogue groups of four warps; a group takes whole windows
constexpr int kSW = 16, kValid = 14;      // window pitch; outputs per window row / column (16 x 16 input pixels)
constexpr int kXchBytes = 8192;       // per group: X0[3][16][16] f32, X2[3][16][16] f32, P0[16][16], C7[16][16]
constexpr uint32_t kBlk = 1536;       // one (dy, k-step) weight block: 48 rows x 16 K x 2 B

struct Win {
  int n, y0, x0, ty, tx;
};

__device__ __forceinline__ int fast_div(int t, unsigned long long magic) {
  return static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(t)) * magic) >> 40);
}
__device__ __forceinline__ Win decode_win(const DenseParams& p, int t) {
  Win r;
  r.n = fast_div(t, p.magic_img);
  const int rem = t - r.n * p.tiles_per_img;
  r.ty = fast_div(rem, p.magic_row);
  r.tx = rem - r.ty * p.tiles_x;
  r.y0 = r.ty * kValid;
  r.x0 = r.tx * kValid;
  return r;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* ptr, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

static __device__ __noinline__ void flag_timeout9(int layer, int win, unsigned have) {
  printf("climsr_b200: dense-block (N=144) dependency timeout: block %d waits for layer %d window %d (counter %u)\n", blockIdx.x, layer, win, have);
  __trap();
}

// s[j] = raw0[j] of lane-1 + raw1[j] (own) + raw2[j] of lane+1  (horizontal taps dx = 0, 1, 2 of one vertical tap)
__device__ __forceinline__ void dx_sum8(float (&s)[8], const uint32_t (&r0)[8], const uint32_t (&r1)[8], const uint32_t (&r2)[8], int lane, bool noshfl = false) {
  if (noshfl) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]) + __uint_as_float(r2[j]);
    return;
  }
  const int lm = (lane + 31) & 31, lp = (lane + 1) & 31;
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const float2 a = make_float2(__uint_as_float(__shfl_sync(0xffffffffu, r0[j], lm)), __uint_as_float(__shfl_sync(0xffffffffu, r0[j + 1], lm)));
    const float2 c = make_float2(__uint_as_float(__shfl_sync(0xffffffffu, r2[j], lp)), __uint_as_float(__shfl_sync(0xffffffffu, r2[j + 1], lp)));
    const float2 t = __fadd2_rn(__fadd2_rn(a, make_float2(__uint_as_float(r1[j]), __uint_as_float(r1[j + 1]))), c);
    s[j] = t.x;
    s[j + 1] = t.y;
  }
}

__device__ __forceinline__ void st16f(uint32_t addr, const float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    st_shared_v4(addr + q * 16, __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
}
__device__ __forceinline__ void add16f(float (&v)[16], uint32_t addr) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 t = ld_shared_v4(addr + q * 16);
    v[4 * q] += __uint_as_float(t.x); v[4 * q + 1] += __uint_as_float(t.y); v[4 * q + 2] += __uint_as_float(t.z); v[4 * q + 3] += __uint_as_float(t.w);
  }
}

// bias already inside v: activation / gate, bf16, 32 bytes to the concat slice of pixel (n, y, x)
__device__ __forceinline__ void finish_store(const DenseParams& p, int l, float (&v)[16], int n, int y, int x) {
  if (p.act) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {                     // LeakyReLU(0.2) = max(v, 0.2 v)
      const float2 s2 = __fmul2_rn(make_float2(v[j], v[j + 1]), make_float2(0.2f, 0.2f));
      v[j] = fmaxf(v[j], s2.x);
      v[j + 1] = fmaxf(v[j + 1], s2.y);
    }
  }
  const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
  if (p.gate) {                                           // backward form: LeakyReLU derivative of the forward activation
    const uint4* gsrc = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.gate) + pix * p.gate_C + p.L[l].gate_coff);
    const uint4 g0 = __ldg(gsrc), g1 = __ldg(gsrc + 1);
    const uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[2 * i] *= (bf16lo(gw[i]) > 0.f) ? 1.f : p.gate_neg;
      v[2 * i + 1] *= (bf16hi(gw[i]) > 0.f) ? 1.f : p.gate_neg;
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.buf) + pix * p.C + p.L[l].out_coff);
  dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}

}  // namespace

__global__ void __launch_bounds__(kConvThreads, 1)
dense9_block_kernel(const DenseParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.n_slots;
  const uint32_t slots_addr = smem_base;
  const uint32_t w_addr = slots_addr + static_cast<uint32_t>(S) * p.slot_bytes;
  const uint32_t xch_addr = w_addr + 2u * p.wbuf_bytes;
  const uint32_t bias_addr = xch_addr + kGroups * kXchBytes;                   // kDenseMaxLayers x 16 fp32
  const uint32_t bar_addr = bias_addr + kDenseMaxLayers * kNpad * 4;
  // barriers: w_full[2], w_free[2], a_full[S], a_empty[S], acc_full[3], acc_empty[3]
  auto bar_w_full = [&](int b) { return bar_addr + 8u * b; };
  auto bar_w_free = [&](int b) { return bar_addr + 8u * (2 + b); };
  auto bar_a_full = [&](int s) { return bar_addr + 8u * (4 + s); };
  auto bar_a_empty = [&](int s) { return bar_addr + 8u * (4 + S + s); };
  auto bar_acc_full = [&](int b) { return bar_addr + 8u * (4 + 2 * S + b); };
  auto bar_acc_empty = [&](int b) { return bar_addr + 8u * (4 + 2 * S + kAcc9 + b); };
  const uint32_t tmem_slot_addr = bar_addr + 8u * (4 + 2 * S + 2 * kAcc9);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (bias_addr - smem_base));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot_addr - smem_base));
  volatile uint32_t* progress = tmem_slot + 1;             // [2]: ring entries seen full, per issuer
  volatile uint32_t* started = tmem_slot + 3;              // [2]: windows (+1) whose first M tile has passed its accumulator wait, per issuer
  volatile uint32_t* seen = tmem_slot + 5;                 // [3]: per accumulator, uses (+1) whose "full" wait some epilogue warp has passed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nl = p.n_layers;

  griddep_launch_dependents();

  // [k-block][dy][k-step] blocks in global memory -> [k-block][k-step][dy] in shared memory (one thread, 3 * ksteps bulk copies)
  auto load_weights = [&](int l, int b) {
    mbar_arrive_expect_tx(bar_w_full(b), p.L[l].w_bytes);
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.L[l].wpk);
    const uint32_t dst0 = w_addr + b * p.wbuf_bytes;
    const int ksteps = p.L[l].ksteps;
    for (int kb = 0; kb * 4 < ksteps; ++kb) {
      const int ks_here = min(4, ksteps - kb * 4);
      for (int ks = 0; ks < ks_here; ++ks)
        for (int dy = 0; dy < 3; ++dy)
          bulk_load(dst0 + (static_cast<uint32_t>(kb * 4 + ks) * 3u + dy) * kBlk, wsrc + (static_cast<size_t>(kb) * 12 + dy * ks_here + ks) * kBlk, kBlk,
                    bar_w_full(b));
    }
  };

  if (threadIdx.x == 0) {
    progress[0] = 0; progress[1] = 0; started[0] = 0; started[1] = 0; seen[0] = 0; seen[1] = 0; seen[2] = 0;
    tma_prefetch_desc(&tmap);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_w_full(b), 1);
      mbar_init(bar_w_free(b), kMmaWarps);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_a_full(s), 1);
      mbar_init(bar_a_empty(s), 1);
    }
    for (int b = 0; b < kAcc9; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), kEpilogueWarps / kGroups);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
    for (int l = 0; l < 2 && l < nl; ++l) load_weights(l, l);   // constants: before griddepcontrol.wait
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot_addr, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < nl * kNpad; i += blockDim.x) bias_s[i] = p.L[i / kNpad].bias[i % kNpad];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  griddep_wait();
  timeline_start(p.timeline, p.launch_id);

  const int G = static_cast<int>(gridDim.x);
  const int nwin = (p.num_tiles - static_cast<int>(blockIdx.x) + G - 1) / G;   // windows of this CTA in every layer (>= 1)

  if (warp == 0) {
    // ===================== TMA producer =====================
    int slot = 0;
    uint32_t phase = 0;
    for (int l = 0; l < nl; ++l) {
      const unsigned* fl = p.flags + static_cast<size_t>(l > 0 ? l - 1 : 0) * p.num_tiles;
      for (int t = blockIdx.x; t < p.num_tiles; t += G) {
        const Win w = decode_win(p, t);
        if (l > 0 && !(p.dbg & 2)) {
          if (lane < 9) {
            const int dy = lane / 3 - 1, dx = lane - (lane / 3) * 3 - 1;
            const int yy = w.ty + dy, xx = w.tx + dx;
            if (yy >= 0 && yy < p.tiles_y && xx >= 0 && xx < p.tiles_x) {
              const int idx = w.n * p.tiles_per_img + yy * p.tiles_x + xx;
              unsigned have, spins = 0;
              while ((have = ld_acquire_gpu(fl + idx)) < 1u) {
                if (++spins > (1u << 24)) flag_timeout9(l - 1, idx, have);
                __nanosleep(40);
              }
            }
          }
          __syncwarp();
          if (!(p.dbg & 4)) fence_proxy_async_all();      // once per window (rdb_tc.cu)
        }
        for (int kb = 0; kb < p.L[l].n_kblocks; ++kb) {
          mbar_wait_spin(bar_a_empty(slot), phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_a_full(slot), p.win_bytes);
            tma_load_4d(slots_addr + slot * p.slot_bytes, &tmap, bar_a_full(slot), kb * 64, w.x0 - 1, w.y0 - 1, w.n);
          }
          __syncwarp();
          if (++slot == S) { slot = 0; phase ^= 1; }
        }
      }
      if (l >= 1 && l + 1 < nl) {
        const int b = (l + 1) & 1;
        mbar_wait_spin(bar_w_free(b), static_cast<uint32_t>((l - 1) >> 1) & 1u);
        if (elect_one()) load_weights(l + 1, b);
        __syncwarp();
      }
    }
  } else if (warp <= kMmaWarps) {
    // ===================== MMA issuers: alternate windows of the CTA's window sequence q = layer * nwin + it =====================
    const int mw = warp - 1;
    const uint32_t idesc = make_idesc_bf16(kTileM, kN9);
    constexpr uint32_t b_step16 = static_cast<uint32_t>(kN9 * 32) >> 4;       // one k-step of weights (144 rows x 16 K) in 16-byte units
    constexpr uint32_t tile16 = static_cast<uint32_t>(8 * kSW * 128) >> 4;    // eight window rows
    const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);             // SBO = 8 rows x 128 B, version 1, SWIZZLE_128B
    const uint32_t b_hi = (256u >> 4) | (1u << 14);
    const uint32_t a_lbo = (16u >> 4) << 16, b_lbo = (128u >> 4) << 16;
    int slot = 0;
    uint32_t phase = 0;
    int entry = 0;
    int q = 0;
    uint32_t own = 0;
    for (int l = 0; l < nl; ++l) {
      const int nkb = p.L[l].n_kblocks, ksteps = p.L[l].ksteps;
      const uint32_t wl_addr = w_addr + static_cast<uint32_t>(l & 1) * p.wbuf_bytes;
      bool have_w = false;
      for (int it = 0; it < nwin; ++it, ++q) {
        if ((q & 1) != mw) {
          for (int i = 0; i < nkb; ++i, ++entry) {
            own &= ~(1u << (entry & 31));
            if (++slot == S) { slot = 0; phase ^= 1; }
          }
          continue;
        }
        if (!have_w) {
          mbar_wait_spin(bar_w_full(l & 1), static_cast<uint32_t>(l >> 1) & 1u);
          have_w = true;
        }
        for (int kb = 0; kb < nkb; ++kb, ++entry) {
          own |= 1u << (entry & 31);
          const int need = entry - S;
          if (need >= 0 && !((own >> (need & 31)) & 1u)) {
            while (progress[1 - mw] <= static_cast<uint32_t>(need)) {
            }
          }
          mbar_wait_spin(bar_a_full(slot), phase);
          progress[mw] = static_cast<uint32_t>(entry) + 1u;
          tc_fence_after();
          const int ks_here = min(4, ksteps - kb * 4);
          const uint32_t a16 = ((slots_addr + slot * p.slot_bytes) >> 4) | a_lbo;
          const uint32_t b16 = ((wl_addr >> 4) + static_cast<uint32_t>(kb) * 4u * b_step16) | b_lbo;
          const bool last_kb = kb == nkb - 1;
          for (int hh = 0; hh < 2; ++hh) {
            const int m = 2 * q + hh;
            const int use = m / 3, buf = m - use * 3;
            if (kb == 0) {
              // Accumulators are shared by the two issuers (tile m = 2q + hh uses buffer m % 3, use number m / 3), and a parity wait for
              // use u is only unambiguous once the wait of use u-1 has passed.  Use u-1 of this buffer is tile m-3: for hh = 0 that is
              // this issuer's own previous window (program order), for hh = 1 the FIRST tile of window q-1 - the other issuer's.
              if (hh == 1 && q >= 1) {
                while (started[1 - mw] < static_cast<uint32_t>(q)) {
                }
              }
              mbar_wait_spin(bar_acc_empty(buf), (static_cast<uint32_t>(use) & 1u) ^ 1u);
              tc_fence_after();
              if (hh == 0) started[mw] = static_cast<uint32_t>(q) + 1u;
            }
            if (elect_one()) {
              const uint32_t dh = tmem_base + static_cast<uint32_t>(buf * kAccStride);
              const uint32_t ah = a16 + static_cast<uint32_t>(hh) * tile16;
              uint32_t bh = b16;
              for (int ks = 0; ks < ks_here; ++ks, bh += b_step16) umma_bf16_split(dh, ah + ks * 2, a_hi, bh, b_hi, idesc, (kb | ks) ? 1u : 0u);
              if (last_kb) umma_commit(bar_acc_full(buf));
              if (hh == 1) umma_commit(bar_a_empty(slot));
            }
            __syncwarp();
          }
          if (++slot == S) { slot = 0; phase ^= 1; }
        }
      }
      if (elect_one()) umma_commit(bar_w_free(l & 1));
      __syncwarp();
    }
  } else {
    // ===================== epilogue: four groups of four warps; group g takes the windows q = l * nwin + it with q % 4 == g ==========
    const int ew = warp - 1 - kMmaWarps;
    const int g = ew >> 2;
    const int lg = warp & 3;                              // TMEM lane quadrant (hardware rule: warp % 4) = tile rows 2 lg, 2 lg + 1
    const int half = lane >> 4;                           // 0: tile row 2 lg, 1: tile row 2 lg + 1
    const int c = lane & 15;                              // window column
    const bool col_ok = (c >= 1) && (c <= kValid);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t xg = xch_addr + static_cast<uint32_t>(g) * kXchBytes;
    // X0[j] (j = 0..2): T0 of tile row 2j+1, read by row 2j+2.  X2[j] (j = 0..2): T2 of tile row 2j+2, read by row 2j+1.
    const uint32_t x0_addr = xg, x2_addr = xg + 3072, p0_addr = xg + 6144, c7_addr = xg + 7168;
    const uint32_t col_off = static_cast<uint32_t>(c) * 64u;
    int l = 0, it = g;
    while (it >= nwin && l < nl) { it -= nwin; ++l; }
    for (int q = g; q < nwin * nl; q += kGroups) {
      const int t = blockIdx.x + it * G;
      const Win w = decode_win(p, t);
      const int x = w.x0 - 1 + c;
      const bool x_ok = col_ok && x < p.W;
      for (int hh = 0; hh < 2; ++hh) {
        const int m = 2 * q + hh;
        const int use = m / 3, buf = m - use * 3;
        const uint32_t t_addr = t_lane + static_cast<uint32_t>(buf * kAccStride);
        const int wr = hh * 8 + lg * 2 + half;            // window row of this thread's position; output row y = y0 - 1 + wr
        const int y = w.y0 - 1 + wr;
        float acc[16];
        float tt[16];
        uint32_t raw[2][3][8];
        // The accumulators rotate over the epilogue groups, so this group may be two uses ahead of the barrier: a parity wait for use u
        // is only unambiguous once use u-1 has been seen full by its own group (which then publishes it in `seen`).
        if (use >= 1) {
          while (seen[buf] < static_cast<uint32_t>(use)) {
          }
        }
        mbar_wait(bar_acc_full(buf), static_cast<uint32_t>(use) & 1u);   // bounded: a protocol bug traps here instead of hanging
        if (lane == 0) seen[buf] = static_cast<uint32_t>(use) + 1u;
        if (p.dbg & 8) {                                   // timing experiment: MMA pipeline only
          tc_fence_after();
          tmem_ld8(t_addr, raw[0][0]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty(buf));
          continue;
        }
        tc_fence_after();
        // ---- vertical tap 0: T0(row) is a term of the row BELOW
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) tmem_ld8(t_addr + dx * kNpad + jj * 8, raw[jj][dx]);
        tmem_ld_wait();
        {
          float s0[8], s1[8];
          dx_sum8(s0, raw[0][0], raw[0][1], raw[0][2], lane, p.dbg & 16);
          dx_sum8(s1, raw[1][0], raw[1][1], raw[1][2], lane, p.dbg & 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) { tt[j] = s0[j]; tt[8 + j] = s1[j]; }
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) tmem_ld8(t_addr + 2 * 3 * kNpad + dx * kNpad + jj * 8, raw[jj][dx]);   // vertical tap 2 in flight
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float other = __shfl_xor_sync(0xffffffffu, tt[j], 16);
          acc[j] = bias_s[l * kNpad + j] + (half ? other : 0.f);             // row 2lg+1 takes T0 of row 2lg
        }
        if (half) {                                        // T0 of row 2lg+1 -> row 2lg+2: next warp, or the first row of the next M tile
          if (lg < 3) st16f(x0_addr + static_cast<uint32_t>(lg) * 1024u + col_off, tt);
          else if (hh == 0) st16f(p0_addr + col_off, tt);
        }
        tmem_ld_wait();
        // ---- vertical tap 2: T2(row) is a term of the row ABOVE
        {
          float s0[8], s1[8];
          dx_sum8(s0, raw[0][0], raw[0][1], raw[0][2], lane, p.dbg & 16);
          dx_sum8(s1, raw[1][0], raw[1][1], raw[1][2], lane, p.dbg & 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) { tt[j] = s0[j]; tt[8 + j] = s1[j]; }
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) tmem_ld8(t_addr + 3 * kNpad + dx * kNpad + jj * 8, raw[jj][dx]);       // vertical tap 1 in flight
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float other = __shfl_xor_sync(0xffffffffu, tt[j], 16);
          acc[j] += half ? 0.f : other;                    // row 2lg takes T2 of row 2lg+1
        }
        if (!half) {
          if (lg > 0) {
            st16f(x2_addr + static_cast<uint32_t>(lg - 1) * 1024u + col_off, tt);   // T2 of row 2lg -> row 2lg-1: previous warp
          } else if (hh == 1) {
            // first row of the second M tile: completes the LAST row of the first tile (window row 7), parked in C7
            add16f(tt, c7_addr + col_off);
            if (x_ok && y - 1 < p.H) finish_store(p, l, tt, w.n, y - 1, x);
          }
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty(buf));    // everything of this accumulator is in registers
        // ---- vertical tap 1: own row
        {
          float s0[8], s1[8];
          dx_sum8(s0, raw[0][0], raw[0][1], raw[0][2], lane, p.dbg & 16);
          dx_sum8(s1, raw[1][0], raw[1][1], raw[1][2], lane, p.dbg & 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[j] += s0[j]; acc[8 + j] += s1[j]; }
        }
        if (half && lg == 3 && hh == 0) st16f(c7_addr + col_off, acc);       // window row 7: waits for T2 of row 8 (next M tile)
        if (!(p.dbg & 64)) named_bar_sync(1 + g, 128);     // X0 / X2 / P0 / C7 of this tile are written
        bool row_ok = true;
        if (!half) {                                       // needs T0 of the row above
          if (lg > 0) add16f(acc, x0_addr + static_cast<uint32_t>(lg - 1) * 1024u + col_off);
          else if (hh == 1) add16f(acc, p0_addr + col_off);
          else row_ok = false;                             // window row 0: halo
        } else {                                           // needs T2 of the row below
          if (lg < 3) add16f(acc, x2_addr + static_cast<uint32_t>(lg) * 1024u + col_off);
          else row_ok = false;                             // window row 7: finished by the next M tile; window row 15: halo
        }
        if (row_ok && x_ok && y < p.H && !(p.dbg & 32)) finish_store(p, l, acc, w.n, y, x);
        if (!(p.dbg & 64)) named_bar_sync(1 + g, 128);     // exchange buffers free again; all stores of the tile issued
      }
      if (ew == g * 4 && lane == 0) {
        if (!(p.dbg & 1)) __threadfence();                 // cumulative: the group's stores, ordered before this by the barrier
        red_release_gpu_add(p.flags + static_cast<size_t>(l) * p.num_tiles + t, 1u);
      }
      it += kGroups;
      while (it >= nwin && l < nl) { it -= nwin; ++l; }
    }
  }

  tc_fence_before();
  __syncthreads();
  timeline_end(p.timeline, p.launch_id);
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

size_t dense9_smem_bytes(const DenseParams& p) {
  return 1024 + static_cast<size_t>(p.n_slots) * p.slot_bytes + 2u * p.wbuf_bytes + kGroups * kXchBytes + kDenseMaxLayers * kNpad * 4 +
         8 * (4 + 2 * p.n_slots + 2 * kAcc9) + 32;
}

int launch_dense9_block(const DenseParams& p, const CUtensorMap& tmap, int num_sms, cudaStream_t stream) {
  const size_t smem = dense9_smem_bytes(p);
  if (smem > static_cast<size_t>(kSmemLimit) || p.n_layers < 1 || p.n_layers > kDenseMaxLayers || p.n_slots < 2 || p.n_slots > 8)
    return static_cast<int>(cudaErrorInvalidValue);
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return static_cast<int>(cudaErrorInvalidDevice);
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(dense9_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[dev] = true;
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
  return static_cast<int>(cudaLaunchKernelEx(&cfg, dense9_block_kernel, p, tmap));
}

}  // namespace csr
#endif  // CSR_EXPERIMENTS
