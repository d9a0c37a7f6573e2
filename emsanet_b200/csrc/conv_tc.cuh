// Parameter blocks shared by the tcgen05 convolution kernels and the C-ABI glue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace eb {

constexpr int kMaxTaps = 9;

enum ConvFlags : uint32_t {
  kBias = 1u << 0,      // + bias[c] (fp32) on the accumulator
  kRelu = 1u << 1,      // max(.,0): after the residual add when kAuxAdd is set, else on the accumulator
  kAuxAdd = 1u << 2,    // + aux (bf16 tensor addressed like the output)
  kAuxMask = 1u << 3,   // value kept only where aux > 0 (ReLU backward)
  kStats = 1u << 4,     // per-channel sum / sum of squares of the stored values -> stats[0:C], stats[C:2C]
  kStatsSum = 1u << 5,  // with kStats: only the sums, stats is [C] (bias gradient accumulated straight into .grad)
  kBnBwd = 1u << 6,     // with kAuxMask|kStats: aux = raw BN input x; mask = x*bn_scale+bn_shift > 0; stats = (sum g, sum g*x)
};

// Implicit-GEMM convolution, stride 1 over up to two "views" of the input (parity views implement stride 2):
//   out[n,h,w,co] = sum_t sum_ci  in_view[t][n, h+dy[t], w+dx[t], ci] * wgt[t][co][ci]
// A tiles are [128 pixels x 64 ci] TMA boxes (box = 64 x bw x bh x bn, out-of-bounds = zero padding),
// B tiles are [block_n co x 64 ci] TMA boxes of the re-laid-out weights [taps][Cout_pad][Cin].
struct ConvParams {
  CUtensorMap map_a[2];
  CUtensorMap map_b;
  int N, H, W;             // output extent (pixels)
  int Cin, Cout;           // Cin % 64 == 0; Cout % 8 == 0 (real output channels)
  int block_n;             // UMMA N (multiple of 16, <= 256); weights padded to a multiple of block_n rows
  int taps;
  int tap_view[kMaxTaps], tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_w[kMaxTaps];
  int lbw, lbh, lbn;       // log2 of the pixel box (bw*bh*bn == 128)
  int tiles_w, tiles_h, tiles_n, tiles_c;   // tile grid (tiles_c = ceil(Cout / block_n))
  int stages;
  int ksub;                // 64-channel (tap, K-block) sub-blocks per pipeline stage: one barrier round trip and one
                           // tcgen05.commit per stage amortise over 4*ksub MMAs (the single issuing thread is the
                           // bottleneck for narrow tiles)
  int b_resident;          // 1: the whole weight tensor of this channel tile stays in smem (loaded once per CTA)
  long long* dbg_buf;      // experiments only (debug & 8): clock64 trace of CTA 0
  int debug;               // experiments only: 1 = skip global stores, 2 = skip MMA issue, 4 = skip TMA loads
  int cluster;             // CTAs per cluster (1, 2 or 4): consecutive pixel tiles share the weight tile by TMA multicast
  uint32_t flags;
  __nv_bfloat16* out;      // element strides below; channel c of pixel at out + off + c
  long long out_sn, out_sh, out_sw;
  const __nv_bfloat16* aux;
  long long aux_sn, aux_sh, aux_sw;
  const float* bias;
  float* stats;            // [2][Cout] fp32, atomically accumulated
  const float* bn_scale;   // kBnBwd
  const float* bn_shift;
};

// Weight gradient: dW[t][co][ci] += sum_pixels dY[n,h,w,co] * X_view[t][n, h+dy[t], w+dx[t], ci]
// GEMM with M = co (128 per CTA), N = ci (block_n per CTA), K = pixels (64 per stage), both operands MN-major.
struct WgradParams {
  CUtensorMap map_dy;      // box 64co x bw x bh x bn  (64 pixels)
  CUtensorMap map_x[2];    // box 64ci x bw x bh x bn
  int N, H, W;             // extent of dY
  int Cin, Cout;
  int block_n;             // ci per CTA (multiple of 64, taps_per_item * block_n <= 512)
  int taps;                // taps handled per work item (accumulators in TMEM)
  int tap_groups;          // work items along taps
  int tap_view[kMaxTaps], tap_dy[kMaxTaps], tap_dx[kMaxTaps];  // indexed by absolute tap id
  int lbw, lbh, lbn;       // pixel box, bw*bh*bn == 64
  int tiles_w, tiles_h, tiles_n;
  int ksplit;              // CTAs along the pixel reduction
  int stages;
  float* dw;               // fp32, element (t, co, ci) at dw + co*dw_sco + ci*dw_sci + t*dw_st  (atomic add)
  long long dw_sco, dw_sci, dw_st;
  int total_taps;
  int vec_ok;              // dw is 16-byte aligned: vector reductions allowed
  int bulk;                // 0: scalar atomics; 1 / 2 / 3: TMA bulk reduce-add of whole rows (1x1 / 3-tap / 3x3 via ws)
  float* ws;               // bulk == 3: fp32 [3][Cout][Cin][3] staging (zero on entry, re-zeroed by the finish kernel)
};

}  // namespace eb
