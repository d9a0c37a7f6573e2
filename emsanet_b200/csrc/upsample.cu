// Learned upsampling 'learned-3x3-zeropad' (MT/model/upsampling.py:39-96): nearest x2, then a depthwise 3x3 with zero
// padding on the UPSAMPLED map, + bias — forward, input gradient and weight/bias gradient (autograd of the same,
// entered at main.py:598).  The upsampled tensor is never materialised.
//
// These are bandwidth kernels with a 3x3 (forward) / 4x4 (input gradient) neighbourhood: read straight from global
// memory every operand is fetched 9x / 4x through L1, and most of a thread's loads in flight are redundant.  All
// three kernels therefore stage a tile (+ halo) of the small operand in shared memory with coalesced 16-byte loads,
// each element exactly once, and take the neighbourhood from there; channels are the fastest thread index so every
// global access of a warp is contiguous.  weights fp32 [Creal][9] (reference layout [C,1,3,3]).
//
// Source-pixel-centric mapping (forward / weight gradient): source pixel (h, w) with its 3x3 source neighbourhood S
// (zero outside the map == zero padding of the upsampled map) determines the 2x2 output block (2h+a, 2w+b):
//   out(a,b) = bias + sum_{ky,kx} w[ky][kx] * S[ry(a,ky)][rx(b,kx)],   ry(0,.) = (0,1,1), ry(1,.) = (1,1,2)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/emsanet_b200.h"
#include "common.h"

namespace eb {

__device__ __forceinline__ int up_r2(int a, int k) { return a == 0 ? (k == 0 ? 0 : 1) : (k == 2 ? 2 : 1); }

__device__ __forceinline__ void cvt8u(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return u;
}

struct UpTile {
  int N, H, W, C, Creal;   // source extent, channel pitch, real channels
  int CB;                  // channels per block (multiple of 8, <= 128)
  int TH, TW;              // source pixels per tile
  int tiles_h, tiles_w, chunks;
};

// stage rows [r0, r0+rows) x cols [c0, c0+cols) x channels [cb0, cb0+CB) of a [N][Himg][Wimg][C] tensor in smem as
// [rows][cols][CB] (zero outside the image).  cp.async (LDGSTS, 16 bytes, zero-fill for out-of-image elements): every
// load of the tile is in flight at once — a load -> st.shared loop would serialise on the global-memory latency.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// Per-thread map of the (<= 4) 16-byte vectors it copies of every tile row: column, channel group — computed ONCE per
// kernel (the only divisions of the staging path), reused for every row of every tile.
struct RowMap {
  int n;           // vectors of a row handled by this thread
  int v[4];        // vector index within the row ([cols][CB8] order)
  int cc[4];       // tile column
  int c8x8[4];     // channel offset (elements)
};
__device__ __forceinline__ RowMap make_row_map(int cols, int CB8) {
  RowMap m;
  m.n = 0;
  const int row_vecs = cols * CB8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int v = threadIdx.x + k * 256;
    m.v[k] = v; m.cc[k] = 0; m.c8x8[k] = 0;
    if (v < row_vecs) {
      m.cc[k] = v / CB8;
      m.c8x8[k] = (v - m.cc[k] * CB8) * 8;
      m.n = k + 1;
    }
  }
  return m;
}
__device__ __forceinline__ void stage_tile(const __nv_bfloat16* __restrict__ src, uint4* __restrict__ dst, int n, int Himg,
                                           int Wimg, int C, int cb0, const RowMap& m, int row_vecs, int r0, int rows,
                                           int c0) {
  const uint32_t dst0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  for (int rr = 0; rr < rows; ++rr) {
    const int r = r0 + rr;
    const bool row_ok = r >= 0 && r < Himg;
    const __nv_bfloat16* row = src + ((static_cast<size_t>(n) * Himg + (row_ok ? r : 0)) * Wimg) * C + cb0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < m.n) {
        const int c = c0 + m.cc[k];
        const bool ok = row_ok && c >= 0 && c < Wimg;
        cp_async16(dst0 + (rr * row_vecs + m.v[k]) * 16,
                   ok ? static_cast<const void*>(row + static_cast<size_t>(c) * C + m.c8x8[k])
                      : static_cast<const void*>(src), ok ? 16u : 0u);
      }
    }
  }
}

// Persistent tile walk shared by the three kernels: a block owns one channel chunk and walks tiles of it with a
// two-deep cp.async pipeline (tile i+1 is in flight while tile i is consumed); per-thread weights are set up once.
struct TileWalk {
  int chunk, bic, bpc;          // channel chunk, block index within the chunk, blocks per chunk
  int tiles_per_img, total_tiles;
};
__device__ __forceinline__ TileWalk tile_walk(const UpTile& t, int blocks_per_chunk) {
  TileWalk w;
  w.bpc = blocks_per_chunk;
  w.chunk = blockIdx.x / blocks_per_chunk;
  w.bic = blockIdx.x - w.chunk * blocks_per_chunk;
  w.tiles_per_img = t.tiles_h * t.tiles_w;
  w.total_tiles = w.tiles_per_img * t.N;
  return w;
}
__device__ __forceinline__ void tile_coords(const UpTile& t, const TileWalk& w, int tl, int& n, int& h0, int& w0) {
  n = tl / w.tiles_per_img;
  const int rem = tl - n * w.tiles_per_img;
  const int th = rem / t.tiles_w;
  h0 = th * t.TH;
  w0 = (rem - th * t.tiles_w) * t.TW;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------ forward
// Output-pixel-centric with combined weights: because the 3x3 filter runs over a nearest-upsampled map, output
// (2h+a, 2w+b) only sees the 2x2 source patch rows h-1+a.., cols w-1+b.., each source pixel weighted by the SUM of the
// taps that land on it:  out = bias + sum_{r,s in 0..1} CW[r][s] * S[h-1+a+r][w-1+b+s],
//   CW[r][s] = sum_{ky in G(a,r)} sum_{kx in G(b,s)} w[ky][kx],  G(0,0)={0} G(0,1)={1,2} G(1,0)={0,1} G(1,1)={2}
// 4 FMAs per output element instead of 9.  A thread keeps one parity (a, b) and 8 channels: 32 combined weights.
__device__ __forceinline__ bool up_in_group(int par, int r, int k) {
  return par == 0 ? (r == 0 ? k == 0 : k >= 1) : (r == 0 ? k <= 1 : k == 2);
}

__global__ void __launch_bounds__(256, 2) upsample_dw_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                 const float* __restrict__ wgt,
                                                                 const float* __restrict__ bias,
                                                                 __nv_bfloat16* __restrict__ y, UpTile t,
                                                                 int blocks_per_chunk) {
  extern __shared__ uint4 tile[];   // 2 x [TH+2][TW+2][CB8]
  const int CB8 = t.CB >> 3;
  const TileWalk wk = tile_walk(t, blocks_per_chunk);
  const int cb0 = wk.chunk * t.CB;
  const int cols = t.TW + 2;
  const int buf_vecs = (t.TH + 2) * cols * CB8;
  const RowMap rmap = make_row_map(cols, CB8);
  const int a = threadIdx.x >> 7;                    // output row parity of this thread
  const int tl = threadIdx.x & 127;
  const int c8 = tl % CB8;
  const int slot = tl / CB8;
  const int nslots = (128 / CB8) & ~1;               // even: the column parity b of a thread never changes
  const int bpar = slot & 1;
  int n, h0, w0;
  int cur = wk.bic;
  if (cur < wk.total_tiles) {
    tile_coords(t, wk, cur, n, h0, w0);
    stage_tile(x, tile, n, t.H, t.W, t.C, cb0, rmap, cols * CB8, h0 - 1, t.TH + 2, w0 - 1);
  }
  cp_async_commit();
  float cw[2][2][8], bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb0 + c8 * 8 + j;
    bv[j] = c < t.Creal ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float acc = 0.f;
        if (c < t.Creal) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              if (up_in_group(a, r, ky) && up_in_group(bpar, q, kx)) acc += __ldg(wgt + c * 9 + ky * 3 + kx);
        }
        cw[r][q][j] = acc;
      }
  }
  const int Wo = 2 * t.W;
  int buf = 0;
  for (; cur < wk.total_tiles; cur += wk.bpc, buf ^= 1) {
    const int nxt = cur + wk.bpc;
    if (nxt < wk.total_tiles) {
      int n2, h2, w2;
      tile_coords(t, wk, nxt, n2, h2, w2);
      stage_tile(x, tile + (buf ^ 1) * buf_vecs, n2, t.H, t.W, t.C, cb0, rmap, cols * CB8, h2 - 1, t.TH + 2, w2 - 1);
    }
    cp_async_commit();
    cp_async_wait<1>();          // everything but the group just committed has landed: tile `cur` is in smem
    __syncthreads();
    tile_coords(t, wk, cur, n, h0, w0);
    const uint4* tb = tile + buf * buf_vecs;
    if (slot < nslots) {
      for (int ph = 0; ph < t.TH; ++ph) {
        const int h = h0 + ph;
        if (h >= t.H) break;
        for (int xl = slot; xl < 2 * t.TW; xl += nslots) {
          const int pw = xl >> 1;
          if (w0 + pw >= t.W) break;
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = bv[j];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float sv[8];
              cvt8u(tb[((ph + a + r) * cols + pw + bpar + q) * CB8 + c8], sv);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = fmaf(sv[j], cw[r][q][j], o[j]);
            }
          *reinterpret_cast<uint4*>(y + ((static_cast<size_t>(n) * 2 * t.H + 2 * h + a) * Wo + 2 * w0 + xl) * t.C + cb0 +
                                    c8 * 8) = pack8(o);
        }
      }
    }
    __syncthreads();             // the buffer is re-staged two iterations from now
  }
}

// ------------------------------------------------------------------------------------------------ input gradient
// dx[n,h,w,c] = sum_{r,q in 0..3} cw[r][q][c] * dy[2h-1+r][2w-1+q],  cw[r][q] = sum of w[ky][kx] over the (a,ky), (b,kx)
// with a - ky + 2 == r, b - kx + 2 == q.  4 channels per thread (64 combined weights in registers).
__global__ void __launch_bounds__(256, 2) upsample_dw_bwd_input_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                       const float* __restrict__ wgt,
                                                                       __nv_bfloat16* __restrict__ dx, UpTile t,
                                                                       int blocks_per_chunk) {
  extern __shared__ uint4 tile[];   // 2 x [2TH+2][2TW+2][CB8]
  const int CB8 = t.CB >> 3, CB4 = t.CB >> 2;
  const TileWalk wk = tile_walk(t, blocks_per_chunk);
  const int cb0 = wk.chunk * t.CB;
  const int cols = 2 * t.TW + 2, rows = 2 * t.TH + 2;
  const int buf_vecs = rows * cols * CB8;
  const RowMap rmap = make_row_map(cols, CB8);
  const int c4 = threadIdx.x % CB4;
  const int slot = threadIdx.x / CB4, nslots = blockDim.x / CB4;
  const int lgTW = 31 - __clz(t.TW);
  int n, h0, w0;
  int cur = wk.bic;
  if (cur < wk.total_tiles) {
    tile_coords(t, wk, cur, n, h0, w0);
    stage_tile(dy, tile, n, 2 * t.H, 2 * t.W, t.C, cb0, rmap, cols * CB8, 2 * h0 - 1, rows, 2 * w0 - 1);
  }
  cp_async_commit();
  float cw[4][4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) cw[r][q][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cb0 + c4 * 4 + j;
    if (c < t.Creal) {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int bb = 0; bb < 2; ++bb)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) cw[a - ky + 2][bb - kx + 2][j] += __ldg(wgt + c * 9 + ky * 3 + kx);
    }
  }
  int buf = 0;
  for (; cur < wk.total_tiles; cur += wk.bpc, buf ^= 1) {
    const int nxt = cur + wk.bpc;
    if (nxt < wk.total_tiles) {
      int n2, h2, w2;
      tile_coords(t, wk, nxt, n2, h2, w2);
      stage_tile(dy, tile + (buf ^ 1) * buf_vecs, n2, 2 * t.H, 2 * t.W, t.C, cb0, rmap, cols * CB8, 2 * h2 - 1, rows, 2 * w2 - 1);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    tile_coords(t, wk, cur, n, h0, w0);
    const uint2* tile2 = reinterpret_cast<const uint2*>(tile + buf * buf_vecs);   // 4-channel granules: [..][CB4]
    if (slot < nslots) {
      for (int p = slot; p < t.TH * t.TW; p += nslots) {
        const int pw = p & (t.TW - 1), ph = p >> lgTW;      // TW is a power of two
        const int h = h0 + ph, w = w0 + pw;
        if (h >= t.H || w >= t.W) continue;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint2 u = tile2[((2 * ph + r) * cols + 2 * pw + q) * CB4 + c4];
            const float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            acc[0] = fmaf(g0.x, cw[r][q][0], acc[0]);
            acc[1] = fmaf(g0.y, cw[r][q][1], acc[1]);
            acc[2] = fmaf(g1.x, cw[r][q][2], acc[2]);
            acc[3] = fmaf(g1.y, cw[r][q][3], acc[3]);
          }
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(acc[0], acc[1]);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(acc[2], acc[3]);
        *reinterpret_cast<uint2*>(dx + ((static_cast<size_t>(n) * t.H + h) * t.W + w) * t.C + cb0 + c4 * 4) = o;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ weight / bias gradient
// dw[c][k] += sum dy[n,Y,X,c] * up[n,Y+ky-1,X+kx-1,c] ; db[c] += sum dy.  Both the x tile (+halo) and the matching
// 2TH x 2TW dy tile are staged (double-buffered); a thread owns 4 channels and keeps its 10 x 4 sums in registers
// across all tiles of the block.
__global__ void __launch_bounds__(256, 2) upsample_dw_bwd_weight_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                        const __nv_bfloat16* __restrict__ x,
                                                                        float* __restrict__ dw, float* __restrict__ db,
                                                                        UpTile t, int blocks_per_chunk) {
  extern __shared__ uint4 tile[];   // 2 x { x: [TH+2][TW+2][CB8] | dy: [2TH][2TW][CB8] }; float[10][CB] at the end
  const int CB8 = t.CB >> 3, CB4 = t.CB >> 2;
  const TileWalk wk = tile_walk(t, blocks_per_chunk);
  const int cb0 = wk.chunk * t.CB;
  const int cols = t.TW + 2, gcols = 2 * t.TW;
  const int x_vecs = (t.TH + 2) * cols * CB8;
  const int buf_vecs = x_vecs + 2 * t.TH * gcols * CB8;
  const RowMap rmap = make_row_map(cols, CB8), gmap = make_row_map(gcols, CB8);
  const int c4 = threadIdx.x % CB4;
  const int slot = threadIdx.x / CB4, nslots = blockDim.x / CB4;
  const int lgTW = 31 - __clz(t.TW);
  float acc[10][4];
#pragma unroll
  for (int k = 0; k < 10; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
  int n, h0, w0;
  int cur = wk.bic;
  if (cur < wk.total_tiles) {
    tile_coords(t, wk, cur, n, h0, w0);
    stage_tile(x, tile, n, t.H, t.W, t.C, cb0, rmap, cols * CB8, h0 - 1, t.TH + 2, w0 - 1);
    stage_tile(dy, tile + x_vecs, n, 2 * t.H, 2 * t.W, t.C, cb0, gmap, gcols * CB8, 2 * h0, 2 * t.TH, 2 * w0);
  }
  cp_async_commit();
  int buf = 0;
  for (; cur < wk.total_tiles; cur += wk.bpc, buf ^= 1) {
    const int nxt = cur + wk.bpc;
    if (nxt < wk.total_tiles) {
      int n2, h2, w2;
      tile_coords(t, wk, nxt, n2, h2, w2);
      uint4* nb = tile + (buf ^ 1) * buf_vecs;
      stage_tile(x, nb, n2, t.H, t.W, t.C, cb0, rmap, cols * CB8, h2 - 1, t.TH + 2, w2 - 1);
      stage_tile(dy, nb + x_vecs, n2, 2 * t.H, 2 * t.W, t.C, cb0, gmap, gcols * CB8, 2 * h2, 2 * t.TH, 2 * w2);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint2* xt2 = reinterpret_cast<const uint2*>(tile + buf * buf_vecs);
    const uint2* gt2 = reinterpret_cast<const uint2*>(tile + buf * buf_vecs + x_vecs);
    if (slot < nslots) {
      for (int p = slot; p < t.TH * t.TW; p += nslots) {   // out-of-image pixels have zero dy: they add nothing
        const int pw = p & (t.TW - 1), ph = p >> lgTW;      // TW is a power of two
        float g[2][2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const uint2 u = gt2[((2 * ph + a) * gcols + 2 * pw + bb) * CB4 + c4];
            const float2 g0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 g1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            g[a][bb][0] = g0.x; g[a][bb][1] = g0.y; g[a][bb][2] = g1.x; g[a][bb][3] = g1.y;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[9][j] += g[a][bb][j];
          }
#pragma unroll
        for (int dyy = 0; dyy < 3; ++dyy)
#pragma unroll
          for (int dxx = 0; dxx < 3; ++dxx) {
            const uint2 u = xt2[((ph + dyy) * cols + pw + dxx) * CB4 + c4];
            const float2 s0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 s1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            const float sv[4] = {s0.x, s0.y, s1.x, s1.y};
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                if (up_r2(a, ky) != dyy) continue;
#pragma unroll
                for (int bb = 0; bb < 2; ++bb)
#pragma unroll
                  for (int kx = 0; kx < 3; ++kx) {
                    if (up_r2(bb, kx) != dxx) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[ky * 3 + kx][j] = fmaf(g[a][bb][j], sv[j], acc[ky * 3 + kx][j]);
                  }
              }
          }
      }
    }
    __syncthreads();
  }
  // block reduction over the pixel slots, then one atomic per (tap, channel) and block
  cp_async_wait<0>();
  __syncthreads();
  float* red = reinterpret_cast<float*>(tile);   // [10][CB]
  for (int i = threadIdx.x; i < 10 * t.CB; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  if (slot < nslots) {
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&red[k * t.CB + c4 * 4 + j], acc[k][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * t.CB; i += blockDim.x) {
    const int k = i / t.CB, c = cb0 + (i - k * t.CB);
    if (c >= t.Creal) continue;
    if (k < 9) atomicAdd(dw + c * 9 + k, red[i]);
    else atomicAdd(db + c, red[i]);
  }
}

static int fill_tile(UpTile& t, int N, int H, int W, int C, int Creal, int TH, int TW, int cb_max) {
  EB_REQUIRE(C % 8 == 0 && Creal <= C && C >= 8, "learned upsampling: C=%d Creal=%d", C, Creal);
  int CB = C;
  if (CB > cb_max) {
    CB = cb_max;
    while (C % CB) CB -= 8;
  }
  t.N = N; t.H = H; t.W = W; t.C = C; t.Creal = Creal; t.CB = CB;
  t.TH = TH; t.TW = TW;
  t.tiles_h = ceil_div(H, TH); t.tiles_w = ceil_div(W, TW); t.chunks = C / CB;
  return 0;
}

}  // namespace eb

using namespace eb;
#define STREAM static_cast<cudaStream_t>(stream)

static int up_blocks_per_chunk(const UpTile& t, int blocks_per_sm) {
  const int total_tiles = t.N * t.tiles_h * t.tiles_w;
  int bpc = (blocks_per_sm * num_sms()) / t.chunks;
  if (bpc > total_tiles) bpc = total_tiles;
  return bpc < 1 ? 1 : bpc;
}

extern "C" int eb200_upsample_dw_fwd(const void* x, const float* w, const float* b, void* y, int N, int H, int W, int C,
                                     int Creal, void* stream) {
  EB_REQUIRE(x && w && b && y, "eb200_upsample_dw_fwd: null argument");
  UpTile t;
  if (fill_tile(t, N, H, W, C, Creal, 4, C <= 8 ? 128 : 32, 128)) return 1;   // thin tensors: wide tiles (a tile row
                                                                            // must keep all 256 threads busy while staging)
  const int bpc = up_blocks_per_chunk(t, 2);
  const size_t smem = static_cast<size_t>(2) * (t.TH + 2) * (t.TW + 2) * t.CB * 2;
  static bool configured = false;
  if (!configured) {
    EB_CUDA(cudaFuncSetAttribute(upsample_dw_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    configured = true;
  }
  upsample_dw_fwd_kernel<<<bpc * t.chunks, 256, smem, STREAM>>>(
      static_cast<const __nv_bfloat16*>(x), w, b, static_cast<__nv_bfloat16*>(y), t, bpc);
  return launch_check("upsample_dw_fwd_kernel");
}

extern "C" int eb200_upsample_dw_bwd_input(const void* dy, const float* w, void* dx, int N, int H, int W, int C,
                                           int Creal, void* stream) {
  EB_REQUIRE(dy && w && dx, "eb200_upsample_dw_bwd_input: null argument");
  UpTile t;
  if (fill_tile(t, N, H, W, C, Creal, 4, C <= 8 ? 64 : 16, 64)) return 1;
  const int bpc = up_blocks_per_chunk(t, 2);
  const size_t smem = static_cast<size_t>(2) * (2 * t.TH + 2) * (2 * t.TW + 2) * t.CB * 2;
  static bool configured = false;
  if (!configured) {
    EB_CUDA(cudaFuncSetAttribute(upsample_dw_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    configured = true;
  }
  upsample_dw_bwd_input_kernel<<<bpc * t.chunks, 256, smem, STREAM>>>(
      static_cast<const __nv_bfloat16*>(dy), w, static_cast<__nv_bfloat16*>(dx), t, bpc);
  return launch_check("upsample_dw_bwd_input_kernel");
}

extern "C" int eb200_upsample_dw_bwd_weight(const void* dy, const void* x, float* dw, float* db, int N, int H, int W,
                                            int C, int Creal, void* stream) {
  EB_REQUIRE(dy && x && dw && db, "eb200_upsample_dw_bwd_weight: null argument");
  UpTile t;
  if (fill_tile(t, N, H, W, C, Creal, 4, C <= 8 ? 64 : 16, 64)) return 1;
  const int bpc = up_blocks_per_chunk(t, 2);
  size_t smem = static_cast<size_t>(2) * ((t.TH + 2) * (t.TW + 2) + 4 * t.TH * t.TW) * t.CB * 2;
  const size_t red = static_cast<size_t>(10) * t.CB * sizeof(float);
  if (smem < red) smem = red;
  static bool configured = false;
  if (!configured) {
    EB_CUDA(cudaFuncSetAttribute(upsample_dw_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    configured = true;
  }
  upsample_dw_bwd_weight_kernel<<<bpc * t.chunks, 256, smem, STREAM>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), dw, db, t, bpc);
  return launch_check("upsample_dw_bwd_weight_kernel");
}
