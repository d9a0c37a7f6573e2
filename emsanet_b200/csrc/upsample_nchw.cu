// The LAST learned upsampling of a task head fused with the output boundary (MT/model/upsampling.py:85-96 followed by
// the fp32 NCHW tensors EMSANet.forward returns, emsanet/model.py:192-233; semantic head: MT/model/decoder/
// semantic.py:62-75).  At the bench batch the semantic logits are the largest tensor of the step (32 x 40 x 480 x 640):
//
//   before: upsample (read 0.2 GB, write 0.79 GB bf16 NHWC) + layout change (read 0.79, write 1.57 GB fp32 NCHW)
//           backward: layout change (read 1.57, write 0.79) + weight gradient (read 0.79 + 0.2) + input gradient (read 0.79)
//   here  : forward reads the 0.2 GB source and writes the 1.57 GB output once; backward reads the 1.57 GB gradient ONCE
//           and produces dx, dW and db from the same staged tile.
//
// nearest x2 + zero-padded depthwise 3x3 in the combined-weight form of upsample.cu: output (2h+a, 2w+b) sees the 2x2
// source patch rows h-1+a.., cols w-1+b.. with the taps that land on the same source pixel pre-summed (4 FMAs instead
// of 9); the input gradient is the transposed 4x4 stencil over the gradient plane.  Depthwise = channels independent, so
// a warp owns one channel (pair) and its lanes walk along the image row: every global access of a warp is a contiguous
// 128-256 byte run of an NCHW row.  The forward result is rounded to bf16 before it is widened to fp32 — the values the
// network returns are exactly those of the bf16 NHWC activation the unfused path stored.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/emsanet_b200.h"
#include "common.h"

namespace eb {

__device__ __forceinline__ bool upn_in_group(int par, int r, int k) {
  return par == 0 ? (r == 0 ? k == 0 : k >= 1) : (r == 0 ? k <= 1 : k == 2);
}
__device__ __forceinline__ int upn_r2(int a, int k) { return a == 0 ? (k == 0 ? 0 : 1) : (k == 2 ? 2 : 1); }
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int kFTH = 8, kFTW = 32;       // forward: source pixels per tile
constexpr int kFCols = 36;               // smem row pitch in words (34 used)

// grid.x = tiles_w * tiles_h * N, grid.y = C / 8.  256 threads = 8 warps: warp = (channel pair, half of the tile rows).
__global__ void __launch_bounds__(256, 5) upsample_fwd_nchw_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const float* __restrict__ wgt,
                                                                const float* __restrict__ bias, float* __restrict__ y,
                                                                int N, int H, int W, int C, int Creal, int tiles_h,
                                                                int tiles_w) {
  __shared__ uint32_t s[kFTH + 2][4][kFCols];          // [row][channel pair][col]: lanes read consecutive words
  const int tile = blockIdx.x;
  const int n = tile / (tiles_h * tiles_w), rem = tile - n * tiles_h * tiles_w;
  const int h0 = (rem / tiles_w) * kFTH, w0 = (rem % tiles_w) * kFTW;
  const int c0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < (kFTH + 2) * (kFTW + 2); i += 256) {
    const int row = i / (kFTW + 2), col = i - row * (kFTW + 2);
    const int h = h0 - 1 + row, w = w0 - 1 + col;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (h >= 0 && h < H && w >= 0 && w < W)
      v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + h) * W + w) * C + c0));
    s[row][0][col] = v.x; s[row][1][col] = v.y; s[row][2][col] = v.z; s[row][3][col] = v.w;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp & 3, rh = warp >> 2;
  float cw[2][2][2][2][2], bv[2];                     // [channel of the pair][a][b][r][q]
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const int c = c0 + 2 * pair + ch;
    const bool live = c < Creal;
    bv[ch] = live ? __ldg(bias + c) : 0.f;
    float wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = live ? __ldg(wgt + c * 9 + k) : 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
                if (upn_in_group(a, r, ky) && upn_in_group(b, q, kx)) acc += wk[ky * 3 + kx];
            cw[ch][a][b][r][q] = acc;
          }
  }
  __syncthreads();
  const int w = w0 + lane;
  const size_t plane = static_cast<size_t>(2 * H) * (2 * W);
#pragma unroll 1
  for (int pi = 0; pi < kFTH / 2; ++pi) {
    const int ph = rh * (kFTH / 2) + pi;
    const int h = h0 + ph;
    if (h >= H) break;
    float v[2][3][3];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint32_t u = s[ph + dy][pair][lane + dx];
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
        v[0][dy][dx] = f.x;
        v[1][dy][dx] = f.y;
      }
    if (w >= W) continue;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int c = c0 + 2 * pair + ch;
      if (c >= Creal) continue;
      float* dst = y + (static_cast<size_t>(n) * Creal + c) * plane + static_cast<size_t>(2 * h) * (2 * W) + 2 * w;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        float o[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float acc = bv[ch];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 2; ++q) acc = fmaf(cw[ch][a][b][r][q], v[ch][a + r][b + q], acc);
          o[b] = bf16_round(acc);
        }
        *reinterpret_cast<float2*>(dst + static_cast<size_t>(a) * (2 * W)) = make_float2(o[0], o[1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kBTH = 4, kBTW = 32;                  // source pixels per tile
constexpr int kGRows = 2 * kBTH + 2;                // gradient tile rows
constexpr int kGPitch = 72;                         // row pitch in floats: column X lives at index X - 2*w0 + 4
constexpr int kXRows = kBTH + 2, kXCols = kBTW + 2;
constexpr int kGBuf = 8 * kGRows * kGPitch * 4;     // bytes of one gradient buffer (8 channels)
constexpr int kXBuf = kXRows * kXCols * 16;         // bytes of one source buffer ([row][col][8 ch] bf16)
constexpr int kBwdSmem = 2 * (kGBuf + kXBuf) + kBTH * kBTW * 16;

__device__ __forceinline__ void cpa16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa4(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// Persistent: blocks_per_group blocks walk the tiles of one group of 8 channels; warp = channel, lane = source column.
// Tiles are staged with cp.async (no registers held by loads in flight: 11 copies per lane and tile, all outstanding at
// once) into a two-deep ring, so the gradient of tile i+1 streams in while tile i is consumed.  Per source pixel and
// channel: the 4x4 gradient window gives dx (16 FMAs with the combined weights), its inner 2x2 and the 3x3 source
// neighbourhood give the weight gradient (16 FMAs into parity-basis accumulators, folded to the 9 taps at the end) and the
// bias gradient; the sums stay in registers across all tiles of the block and leave through one shuffle reduction + 10
// atomics per warp.  Requires W even (16-byte rows).
__global__ void __launch_bounds__(256, 4) upsample_bwd_nchw_kernel(const float* __restrict__ g,
                                                                const __nv_bfloat16* __restrict__ x,
                                                                const float* __restrict__ wgt,
                                                                __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
                                                                float* __restrict__ db, int N, int H, int W, int C,
                                                                int Creal, int tiles_h, int tiles_w, int blocks_per_group) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t smem_u = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  __nv_bfloat16* dxs = reinterpret_cast<__nv_bfloat16*>(smem + 2 * (kGBuf + kXBuf));   // [kBTH][kBTW][8]
  const int group = blockIdx.x / blocks_per_group, bic = blockIdx.x - group * blocks_per_group;
  const int c0 = group * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = c0 + warp;
  const bool live = c < Creal;
  float cw[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) cw[r][q] = 0.f;
  if (live) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) cw[a - ky + 2][b - kx + 2] += __ldg(wgt + c * 9 + ky * 3 + kx);
  }
  // Weight gradient in a transformed basis: dy at output parity (a, b) only ever meets the 2 x 2 source pixels
  // (a + i, b + j), i, j in {0, 1}, of the 3x3 neighbourhood — 16 products per source pixel instead of 36.  The 9 tap sums
  // are combined from the 16 accumulators once at the end: dW[ky][kx] = sum_{a,b} A[a][b][r2(a,ky) - a][r2(b,kx) - b].
  float A[2][2][2][2], accb = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) A[a][b][i][j] = 0.f;
  const int tiles_per_img = tiles_h * tiles_w;
  const int total_tiles = N * tiles_per_img;
  const int Ho = 2 * H, Wo = 2 * W;

  auto stage = [&](int tile, int buf) {
    const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
    const int h0 = (rem / tiles_w) * kBTH, w0 = (rem % tiles_w) * kBTW;
    const uint32_t gdst = smem_u + buf * (kGBuf + kXBuf) + warp * (kGRows * kGPitch * 4);
    const float* gp = g + (static_cast<size_t>(n) * Creal + (live ? c : 0)) * Ho * Wo;
    if (lane < 18) {
      // lanes 0..15: the 64 aligned columns as 16-byte copies; lane 16 / 17: the left / right halo column
      const int X = lane < 16 ? 2 * w0 + 4 * lane : (lane == 16 ? 2 * w0 - 1 : 2 * w0 + 2 * kBTW);
      const int idx = lane < 16 ? 4 + 4 * lane : (lane == 16 ? 3 : 4 + 2 * kBTW);
      const int avail = X < 0 ? 0 : (Wo - X);                       // floats available from X on
#pragma unroll
      for (int row = 0; row < kGRows; ++row) {
        const int Y = 2 * h0 - 1 + row;
        const bool ok = live && Y >= 0 && Y < Ho && avail > 0;
        const float* src = ok ? gp + static_cast<size_t>(Y) * Wo + X : g;
        const uint32_t d = gdst + (row * kGPitch + idx) * 4;
        if (lane < 16) cpa16(d, src, ok ? static_cast<uint32_t>(min(avail, 4) * 4) : 0u);
        else cpa4(d, src, ok ? 4u : 0u);
      }
    }
    if (threadIdx.x < kXRows * kXCols) {
      const int row = threadIdx.x / kXCols, col = threadIdx.x - row * kXCols;
      const int h = h0 - 1 + row, w = w0 - 1 + col;
      const bool ok = h >= 0 && h < H && w >= 0 && w < W;
      cpa16(smem_u + buf * (kGBuf + kXBuf) + kGBuf + threadIdx.x * 16,
            ok ? static_cast<const void*>(x + ((static_cast<size_t>(n) * H + h) * W + w) * C + c0) : static_cast<const void*>(x),
            ok ? 16u : 0u);
    }
  };

  int tile = bic, buf = 0;
  if (tile < total_tiles) stage(tile, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (; tile < total_tiles; tile += blocks_per_group, buf ^= 1) {
    if (tile + blocks_per_group < total_tiles) stage(tile + blocks_per_group, buf ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float* gs = reinterpret_cast<const float*>(smem + buf * (kGBuf + kXBuf)) + warp * (kGRows * kGPitch);
    const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(smem + buf * (kGBuf + kXBuf) + kGBuf);
#pragma unroll
    for (int ph = 0; ph < kBTH; ++ph) {
      float gw[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float* rowp = gs + (2 * ph + r) * kGPitch + 2 * lane;
        const float2 mid = *reinterpret_cast<const float2*>(rowp + 4);
        gw[r][0] = rowp[3]; gw[r][1] = mid.x; gw[r][2] = mid.y; gw[r][3] = rowp[6];
      }
      float d = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) d = fmaf(cw[r][q], gw[r][q], d);
      dxs[(ph * kBTW + lane) * 8 + warp] = __float2bfloat16_rn(d);
      float xv[3][3];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dxx = 0; dxx < 3; ++dxx) xv[dy][dxx] = __bfloat162float(xs[((ph + dy) * kXCols + lane + dxx) * 8 + warp]);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const float gv = gw[a + 1][b + 1];          // dy at (2h+a, 2w+b)
          accb += gv;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) A[a][b][i][j] = fmaf(gv, xv[a + i][b + j], A[a][b][i][j]);
        }
    }
    __syncthreads();
    if (threadIdx.x < kBTH * kBTW) {
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int h = (rem / tiles_w) * kBTH + (threadIdx.x >> 5), w = (rem % tiles_w) * kBTW + (threadIdx.x & 31);
      if (h < H && w < W)
        *reinterpret_cast<uint4*>(dx + ((static_cast<size_t>(n) * H + h) * W + w) * C + c0) =
            *reinterpret_cast<const uint4*>(dxs + threadIdx.x * 8);
    }   // (the barrier of the next iteration orders these reads before dxs is written again, and the buffer `buf`
        //  is re-staged only in the next iteration's stage() call, after every warp has passed the barrier above)
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float acc[10];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  acc[9] = accb;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) acc[ky * 3 + kx] += A[a][b][upn_r2(a, ky) - a][upn_r2(b, kx) - b];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  if (lane == 0 && live) {
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(dw + c * 9 + k, acc[k]);
    atomicAdd(db + c, acc[9]);
  }
}

}  // namespace eb

using namespace eb;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int eb200_upsample_dw_fwd_nchw(const void* x, const float* w, const float* b, float* y, int N, int H, int W,
                                          int C, int Creal, void* stream) {
  EB_REQUIRE(x && w && b && y && N > 0 && H > 0 && W > 0, "eb200_upsample_dw_fwd_nchw: bad argument");
  EB_REQUIRE(C % 8 == 0 && Creal <= C && Creal > 0, "eb200_upsample_dw_fwd_nchw: C=%d Creal=%d", C, Creal);
  const int tiles_h = ceil_div(H, kFTH), tiles_w = ceil_div(W, kFTW);
  dim3 grid(static_cast<unsigned>(N * tiles_h * tiles_w), static_cast<unsigned>(ceil_div(Creal, 8)));
  upsample_fwd_nchw_kernel<<<grid, 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(x), w, b, y, N, H, W, C, Creal,
                                                     tiles_h, tiles_w);
  return launch_check("upsample_fwd_nchw_kernel");
}

extern "C" int eb200_upsample_dw_bwd_nchw(const float* g, const void* x, const float* w, void* dx, float* dw, float* db,
                                          int N, int H, int W, int C, int Creal, void* stream) {
  EB_REQUIRE(g && x && w && dx && dw && db && N > 0 && H > 0 && W > 0, "eb200_upsample_dw_bwd_nchw: bad argument");
  EB_REQUIRE(C % 8 == 0 && Creal <= C && Creal > 0, "eb200_upsample_dw_bwd_nchw: C=%d Creal=%d", C, Creal);
  const int tiles_h = ceil_div(H, kBTH), tiles_w = ceil_div(W, kBTW);
  const int groups = C / 8, total = N * tiles_h * tiles_w;
  int bpg = (4 * num_sms()) / groups;
  if (bpg > total) bpg = total;
  if (bpg < 1) bpg = 1;
  EB_REQUIRE(W % 2 == 0, "eb200_upsample_dw_bwd_nchw: odd source width %d (rows of the gradient must be 16-byte multiples)", W);
  EB_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "eb200_upsample_dw_bwd_nchw: gradient not 16-byte aligned");
  static bool configured = false;
  if (!configured) {
    EB_CUDA(cudaFuncSetAttribute(upsample_bwd_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    configured = true;
  }
  upsample_bwd_nchw_kernel<<<bpg * groups, 256, kBwdSmem, STREAM>>>(g, static_cast<const __nv_bfloat16*>(x), w,
                                                            static_cast<__nv_bfloat16*>(dx), dw, db, N, H, W, C, Creal,
                                                            tiles_h, tiles_w, bpg);
  return launch_check("upsample_bwd_nchw_kernel");
}
