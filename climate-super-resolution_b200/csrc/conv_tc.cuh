// Implicit-GEMM convolution on tcgen05 tensor cores: parameters shared by kernel and host launcher.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace csr {

constexpr int kConvThreads = 320;     // warp0 TMA producer, warp1 MMA issuer (+TMEM alloc), warps2-9 epilogue
constexpr int kEpilogueThreads = 256;
constexpr int kTileM = 128;           // output pixels (UMMA M) per tile = TH * SW
constexpr int kSmemLimit = 232448;    // 227 KB opt-in dynamic shared memory per CTA on sm_100

struct ConvParams {
  // geometry (input spatial == conv output spatial; stride 1, "same" padding)
  int N, H, W;
  int KH, KW, PH, PW;
  int cin_off;     // first input channel inside the input buffer (multiple of 8)
  int cin;         // input channels used, padded to a multiple of 16 (k-steps = cin/16)
  int npad;        // output channels of this launch padded to a multiple of 16; UMMA N = KW * npad (<= 256)
  int n_store;     // real output channels
  // tiling: a tile is TH rows x SW smem-pitch columns of which the first TW = SW-(KW-1) are real outputs
  int SW, sw_shift, TH, TW;
  int tiles_x, tiles_y, num_tiles;
  int win_rows;    // TH + KH - 1
  int win_bytes;   // win_rows * SW * 128  (== TMA box bytes per 64-channel k-block)
  int slot_bytes;  // win_bytes rounded up to 1024
  int n_kblocks;   // ceil(cin / 64)
  int n_slots;     // activation-window ring depth
  int w_bytes;     // packed weights: KH * (cin/16) * KW * npad * 32
  int tmem_cols;   // power of two >= max(32, 2*KW*npad)
  int use_pdl;     // launch with programmatic stream serialization (prologue overlaps the previous kernel's tail)
  // epilogue
  const float* bias;   // [npad] fp32 (zero padded)
  const void* wpk;     // packed bf16 weights (global), layout [kblock][dy][kstep in kblock][KW*npad/8][2][8][8]
  int act;             // 0 none, 1 leaky-relu 0.2, 2 relu
  float s1, s2;
  const void* r1; int r1_C, r1_coff;
  const void* r2; int r2_C, r2_coff;
  void* out; int out_C, out_coff; int out_mode;   // CsrOutMode
  long long* trace;   // debug: per-role clock64 timestamps of CTA 0 (nullptr = off), [3 roles][64 tiles][4 events]
};

// Returns cudaError_t (as int). tmap: 4-D (C, W, H, N) bf16 tensor map with box (64, SW, win_rows, 1), SWIZZLE_128B.
int launch_conv_tc(const ConvParams& p, const CUtensorMap& tmap, int num_sms, cudaStream_t stream);

}  // namespace csr
