// 3-tap one-dimensional convolution (the 3x1 / 1x3 stride-1 filters of NonBottleneck1D, MT/model/block.py:174-190,
// and their data gradients) on tcgen05 — "halo" formulation.
//
// The 3 taps of a 1-D filter read the same pixels shifted along one image axis.  The generic kernel loads the A tile
// once per tap (3 x 16 KB per 64-channel block); here ONE TMA box with a 2-pixel halo along the tap axis is loaded per
// 64-channel block and the three taps are three UMMA A-descriptors into it:
//
//   tile   : F pixels along the non-tap ("fast") axis x S pixels along the tap ("slow") axis, F*S = 128, one image
//   box    : (64 ch, F, S+2, 1) over a tensor map whose dims are ordered (C, fast, slow, N) — for 1x3 filters the map
//            is the W<->H permuted view of the NHWC tensor, which costs nothing: a pixel's 64 channels are one 128-byte
//            row either way;  out-of-image rows are the TMA zero fill == the filter's zero padding
//   smem   : row r' = s'*F + f  (128-byte swizzled rows); tap with offset o reads rows (o+1)*F .. (o+1)*F+127, i.e. its
//            descriptor starts (o+1)*F*128 bytes into the box — a multiple of the 1024-byte swizzle atom since F >= 8
//
// L2->SM operand traffic per tile drops from 3x to (S+2)/S x the A tile; the weights either stay resident in shared
// memory for the whole CTA (C <= 128) or stream through their own ring.  Control loops are warp-uniform (all lanes
// run them, one lane issues) so that the compiler keeps their state in uniform registers: the single issuing thread
// was the bottleneck of the generic kernel (integer divisions and register->uniform moves between the MMAs).
#pragma once
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace eb {

struct Conv3Params {
  CUtensorMap map_a;
  CUtensorMap map_b;
  int N, ext_f, ext_s;       // image extent along the fast / slow (tap) axis
  int Cout, kblocks;         // Cout % 64 == 0; kblocks = Cin / 64
  int lgF;                   // F = 1 << lgF in {8, 16, 32}; S = 128 >> lgF
  int tiles_f, tiles_s, tiles_c;
  int total_tiles;           // tiles_f * tiles_s * N * tiles_c
  int tap_row[3];            // (offset + 1) * F: first box row of tap t's A operand
  int tap_w[3];              // weight slice of tap t
  int stages_a, stages_b;
  int a_bytes;               // (S + 2) * F * 128
  uint32_t flags;
  __nv_bfloat16* out;
  long long out_sn, out_ss, out_sf;   // element strides: image, slow axis, fast axis
  const __nv_bfloat16* aux;
  long long aux_sn, aux_ss, aux_sf;
  const float* bias;
  float* stats;
  const float* bn_scale;     // kBnBwd: affine of the BatchNorm whose backward is fused into this data gradient
  const float* bn_shift;
};

constexpr int kMaxCout3 = 1024;            // per-CTA statistics / bias scratch (channels)
constexpr int kC3Threads = 320;            // 2 control warps + 8 epilogue warps
constexpr int kC3EpiThreads = 256;
constexpr int kC3Staging = 8 * 2048;       // one private 32x32 bf16 slot per epilogue warp
constexpr int kC3MaxStages = 8;

__host__ __device__ inline int conv3_fixed_smem() { return kC3Staging + 3 * kMaxCout3 * 4 + 512 + 1024; }

// DUAL: one launch runs TWO independent problems of identical geometry (the RGB / depth encoder branches, the semantic /
// instance decoders): even CTAs work on pa, odd CTAs on pb.  At config-2 sizes a wide layer is 2.16 rounds of tiles on 148
// SMs (3 rounds executed); two of them side by side on 74 SMs each are 4.3 rounds (5 executed) and one launch less.
// The CTA index / grid size are read WHERE each warp role starts its tile loop (c3_bx / c3_gd, opaque to the optimiser):
// computed once at the top of the kernel they are hoisted into ordinary registers shared by the three divergent role
// branches, every use in the uniform datapath (TMA / MMA issue) then pays an R2UR, and the launch is 5-12 % slower
// (scripts/conv_single_ab.py on one box: 34.8 vs 31.9 us at C=64, 20.2 vs 18.7 us at C=128; 68 R2UR.BROADCAST in the
// SASS instead of 3).
template <bool DUAL>
__device__ __forceinline__ int c3_bx() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(r));
  return static_cast<int>(DUAL ? r >> 1 : r);
}
template <bool DUAL>
__device__ __forceinline__ int c3_gd() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(r));
  return static_cast<int>(DUAL ? r >> 1 : r);
}

template <int BN, bool RES, uint32_t FLAGS, bool DUAL>
__device__ __forceinline__ void conv3_tc_body(const Conv3Params& p) {
  const uint32_t flags = FLAGS == 0xFFFFFFFFu ? p.flags : FLAGS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  constexpr uint32_t kBTile = BN * 128;                      // one (tap, K-block) weight tile
  const uint32_t a_ring = smem_base;
  const uint32_t b_region = a_ring + p.stages_a * p.a_bytes;
  const uint32_t b_bytes = RES ? 3u * p.kblocks * kBTile : p.stages_b * kBTile;
  const uint32_t stg_base = b_region + b_bytes;
  float* stats_s = reinterpret_cast<float*>(smem + (stg_base - smem_base) + kC3Staging);
  const float* bias_s = stats_s + 2 * kMaxCout3;
  const uint32_t bar_base = stg_base + kC3Staging + 3 * kMaxCout3 * 4;
  auto a_full = [&](uint32_t s) { return bar_base + 8u * s; };
  auto a_empty = [&](uint32_t s) { return bar_base + 8u * (kC3MaxStages + s); };
  auto b_full = [&](uint32_t s) { return bar_base + 8u * (2 * kC3MaxStages + s); };
  auto b_empty = [&](uint32_t s) { return bar_base + 8u * (3 * kC3MaxStages + s); };
  auto tfull_bar = [&](uint32_t a) { return bar_base + 8u * (4 * kC3MaxStages + a); };
  auto tempty_bar = [&](uint32_t a) { return bar_base + 8u * (4 * kC3MaxStages + 2 + a); };
  const uint32_t bres_bar = bar_base + 8u * (4 * kC3MaxStages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (4 * kC3MaxStages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
    for (int s = 0; s < kC3MaxStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kC3EpiThreads);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on
  if (warp >= 2) {
    if (flags & kStats)
      for (int i = threadIdx.x - 64; i < 2 * kMaxCout3; i += kC3EpiThreads) stats_s[i] = 0.f;
    if (flags & kBias)
      for (int i = threadIdx.x - 64; i < kMaxCout3; i += kC3EpiThreads)
        stats_s[2 * kMaxCout3 + i] = i < p.Cout ? __ldg(p.bias + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int F = 1 << p.lgF;
  const int S = 128 >> p.lgF;
  const int tiles_fs = p.tiles_f * p.tiles_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop, lane 0 issues)
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
    if (RES && lane == 0) {
      mbar_arrive_expect_tx(bres_bar, b_bytes);
      for (int t = 0; t < 3; ++t)
        for (int kb = 0; kb < p.kblocks; ++kb)
          tma_load_3d(b_region + (t * p.kblocks + kb) * kBTile, &p.map_b, bres_bar, kb * 64, 0, p.tap_w[t]);
    }
    const int bx = c3_bx<DUAL>(), gd = c3_gd<DUAL>();
    for (int tile = bx; tile < p.total_tiles; tile += gd) {
      const int mt = tile / p.tiles_c, ct = tile - mt * p.tiles_c;
      const int n = mt / tiles_fs, rem = mt - n * tiles_fs;
      const int ts = rem / p.tiles_f, tf = rem - ts * p.tiles_f;
      const int f0 = tf << p.lgF, s0 = ts * S;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(a_empty(sa), pa ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(a_full(sa), p.a_bytes);
          tma_load_4d(a_ring + sa * p.a_bytes, &p.map_a, a_full(sa), kb * 64, f0, s0 - 1, n);
        }
        if (++sa == static_cast<uint32_t>(p.stages_a)) { sa = 0; pa ^= 1u; }
        if (!RES) {
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            mbar_wait(b_empty(sb), pb ^ 1u);
            if (lane == 0) {
              mbar_arrive_expect_tx(b_full(sb), kBTile);
              tma_load_3d(b_region + sb * kBTile, &p.map_b, b_full(sb), kb * 64, ct * BN, p.tap_w[t]);
            }
            if (++sb == static_cast<uint32_t>(p.stages_b)) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, lane 0 issues)
    const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tl = 0;
    if (RES) {
      mbar_wait(bres_bar, 0);
      tc_fence_after();
    }
    const int bx = c3_bx<DUAL>(), gd = c3_gd<DUAL>();
    for (int tile = bx; tile < p.total_tiles; tile += gd, ++tl) {
      const uint32_t acc = tl & 1u;
      mbar_wait(tempty_bar(acc), ((tl >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(a_full(sa), pa);
        tc_fence_after();
        const uint32_t a_base = a_ring + sa * p.a_bytes;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          uint32_t b_addr;
          if (RES) {
            b_addr = b_region + (t * p.kblocks + kb) * kBTile;
          } else {
            mbar_wait(b_full(sb), pb);
            tc_fence_after();
            b_addr = b_region + sb * kBTile;
          }
          if (lane == 0) {
            const uint64_t adesc = make_smem_desc(a_base + p.tap_row[t] * 128, 16, 1024);
            const uint64_t bdesc = make_smem_desc(b_addr, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 4 x UMMA_K(16) = 64 channels; +32 B per step inside the 128 B swizzle row
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | t | k) != 0 ? 1u : 0u);
            if (!RES) umma_commit(b_empty(sb));
          }
          if (!RES) {
            if (++sb == static_cast<uint32_t>(p.stages_b)) { sb = 0; pb ^= 1u; }
          }
        }
        if (lane == 0) umma_commit(a_empty(sa));
        if (++sa == static_cast<uint32_t>(p.stages_a)) { sa = 0; pa ^= 1u; }
      }
      if (lane == 0) umma_commit(tfull_bar(acc));
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 independent warps)
    // Warp (quarter, half) owns accumulator rows 32*quarter..+31 (its TMEM lane quarter) and columns 32*half..+31 of
    // every 64-column chunk; it turns "one row per lane" into 64-byte row segments per lane quad through a private
    // 2 KB smem slot (only __syncwarp needed).  Statistics stay in registers for the whole CTA (the channel tile of a
    // CTA never changes: the grid is a multiple of tiles_c).
    constexpr int NCH = BN / 64;
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t wbuf = stg_base + (warp - 2) * 2048;
    const bool has_bias = flags & kBias;
    const bool relu = flags & kRelu;
    const bool aux_add = flags & kAuxAdd;
    const bool aux_mask = flags & kAuxMask;
    const bool has_aux = aux_add || aux_mask;
    const bool do_stats = flags & kStats;
    const bool sum_only = flags & kStatsSum;
    const bool bn_bwd = flags & kBnBwd;
    const bool relu_in_regs = relu && !aux_add;
    const int piece = lane & 3;
    const int srow = lane >> 2;
    float rsum[NCH][8], rsq[NCH][8];
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) rsum[i][k] = rsq[i][k] = 0.f;
    const int bx = c3_bx<DUAL>(), gd = c3_gd<DUAL>();
    const int ct_fixed = bx % p.tiles_c;
    uint32_t tl = 0;
    for (int tile = bx; tile < p.total_tiles; tile += gd, ++tl) {
      const int mt = tile / p.tiles_c, ct = tile - mt * p.tiles_c;
      const int n = mt / tiles_fs, rem = mt - n * tiles_fs;
      const int ts = rem / p.tiles_f, tf = rem - ts * p.tiles_f;
      const int f0 = tf << p.lgF, s0 = ts * S;
      const uint32_t acc = tl & 1u;
      const int cbase = ct * BN + half * 32 + piece * 8;   // + 64 * chunk = this lane's first channel

      uint32_t ooff[4], aoff[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = quarter * 32 + srow + 8 * j;
        const int f = f0 + (r & (F - 1));
        const int s = s0 + (r >> p.lgF);
        ok[j] = (f < p.ext_f) && (s < p.ext_s);
        ooff[j] = static_cast<uint32_t>(n * p.out_sn + s * p.out_ss + f * p.out_sf) + cbase;
        aoff[j] = has_aux ? static_cast<uint32_t>(n * p.aux_sn + s * p.aux_ss + f * p.aux_sf) + cbase : 0u;
      }
      uint4 av[4], avn[4];
      if (has_aux) {   // chunk 0 of this tile: in flight while the accumulator is still being produced
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[j]) av[j] = __ldg(reinterpret_cast<const uint4*>(p.aux + aoff[j]));
      }
      mbar_wait(tfull_bar(acc), (tl >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256u + half * 32;

#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (has_aux && c + 1 < NCH) {   // next chunk's aux: latency hides behind this chunk
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ok[j]) avn[j] = __ldg(reinterpret_cast<const uint4*>(p.aux + aoff[j] + 64 * (c + 1)));
        }
        {
          uint32_t v[32];
          tmem_ld32(t_row + c * 64, v);
          tmem_ld_wait();
          if (c == NCH - 1) {   // all TMEM reads of this accumulator by this thread are done
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (has_bias) {
            const float4* bp = reinterpret_cast<const float4*>(bias_s + ct * BN + c * 64 + half * 32);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const float4 b4 = bp[q4];
              f[4 * q4 + 0] += b4.x; f[4 * q4 + 1] += b4.y; f[4 * q4 + 2] += b4.z; f[4 * q4 + 3] += b4.w;
            }
          }
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t x0, x1, x2, x3;
            if (relu_in_regs) {
              x0 = pack_bf16x2_relu(f[pc * 8 + 0], f[pc * 8 + 1]); x1 = pack_bf16x2_relu(f[pc * 8 + 2], f[pc * 8 + 3]);
              x2 = pack_bf16x2_relu(f[pc * 8 + 4], f[pc * 8 + 5]); x3 = pack_bf16x2_relu(f[pc * 8 + 6], f[pc * 8 + 7]);
            } else {
              x0 = pack_bf16x2(f[pc * 8 + 0], f[pc * 8 + 1]); x1 = pack_bf16x2(f[pc * 8 + 2], f[pc * 8 + 3]);
              x2 = pack_bf16x2(f[pc * 8 + 4], f[pc * 8 + 5]); x3 = pack_bf16x2(f[pc * 8 + 6], f[pc * 8 + 7]);
            }
            const uint32_t dst = wbuf + lane * 64 + ((pc ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(x0), "r"(x1), "r"(x2), "r"(x3)
                         : "memory");
          }
        }
        __syncwarp();
        uint32_t x[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = srow + 8 * j;
          const uint32_t src = wbuf + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4);
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(x[j][0]), "=r"(x[j][1]), "=r"(x[j][2]), "=r"(x[j][3])
                       : "r"(src));
        }
        float bsc[8], bsh[8];
        if (bn_bwd) {
          const float4* sp = reinterpret_cast<const float4*>(p.bn_scale + cbase + 64 * c);
          const float4* hp = reinterpret_cast<const float4*>(p.bn_shift + cbase + 64 * c);
          const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), h0 = __ldg(hp), h1 = __ldg(hp + 1);
          bsc[0] = s0.x; bsc[1] = s0.y; bsc[2] = s0.z; bsc[3] = s0.w; bsc[4] = s1.x; bsc[5] = s1.y; bsc[6] = s1.z; bsc[7] = s1.w;
          bsh[0] = h0.x; bsh[1] = h0.y; bsh[2] = h0.z; bsh[3] = h0.w; bsh[4] = h1.x; bsh[5] = h1.y; bsh[6] = h1.z; bsh[7] = h1.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!ok[j]) continue;
          if (has_aux) {
            const uint32_t a4[4] = {av[j].x, av[j].y, av[j].z, av[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float2 xv = unpack_bf16x2(x[j][k]);
              const float2 a2 = unpack_bf16x2(a4[k]);
              if (aux_add) {
                xv.x += a2.x; xv.y += a2.y;
                x[j][k] = relu ? pack_bf16x2_relu(xv.x, xv.y) : pack_bf16x2(xv.x, xv.y);
              } else if (bn_bwd) {   // ReLU mask of the BatchNorm output recomputed from its raw input
                xv.x = fmaf(a2.x, bsc[2 * k], bsh[2 * k]) > 0.f ? xv.x : 0.f;
                xv.y = fmaf(a2.y, bsc[2 * k + 1], bsh[2 * k + 1]) > 0.f ? xv.y : 0.f;
                x[j][k] = pack_bf16x2(xv.x, xv.y);
              } else {
                xv.x = a2.x > 0.f ? xv.x : 0.f;
                xv.y = a2.y > 0.f ? xv.y : 0.f;
                x[j][k] = pack_bf16x2(xv.x, xv.y);
              }
            }
          }
          if (do_stats) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 xv = unpack_bf16x2(x[j][k]);
              rsum[c][2 * k] += xv.x;
              rsum[c][2 * k + 1] += xv.y;
              if (bn_bwd) {          // sum g * x (x = aux)
                const uint32_t a4b[4] = {av[j].x, av[j].y, av[j].z, av[j].w};
                const float2 a2 = unpack_bf16x2(a4b[k]);
                rsq[c][2 * k] = fmaf(xv.x, a2.x, rsq[c][2 * k]);
                rsq[c][2 * k + 1] = fmaf(xv.y, a2.y, rsq[c][2 * k + 1]);
              } else if (!sum_only) {
                rsq[c][2 * k] = fmaf(xv.x, xv.x, rsq[c][2 * k]);
                rsq[c][2 * k + 1] = fmaf(xv.y, xv.y, rsq[c][2 * k + 1]);
              }
            }
          }
          *reinterpret_cast<uint4*>(p.out + ooff[j] + 64 * c) = make_uint4(x[j][0], x[j][1], x[j][2], x[j][3]);
        }
        if (has_aux && c + 1 < NCH) {
#pragma unroll
          for (int j = 0; j < 4; ++j) av[j] = avn[j];
        }
        __syncwarp();   // the slot is rewritten by the next chunk
      }
    }
    if (do_stats) {   // one reduction round for the whole CTA
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            rsum[i][k] += __shfl_xor_sync(0xffffffffu, rsum[i][k], off);
            if (!sum_only) rsq[i][k] += __shfl_xor_sync(0xffffffffu, rsq[i][k], off);
          }
        }
        const int gc = ct_fixed * BN + i * 64 + half * 32 + piece * 8;
        if (lane < 4) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            atomicAdd(&stats_s[gc + k], rsum[i][k]);
            if (!sum_only) atomicAdd(&stats_s[kMaxCout3 + gc + k], rsq[i][k]);
          }
        }
      }
      named_bar_sync(1, kC3EpiThreads);
      for (int cidx = threadIdx.x - 64; cidx < BN; cidx += kC3EpiThreads) {
        const int gc = ct_fixed * BN + cidx;
        atomicAdd(p.stats + gc, stats_s[gc]);
        if (!sum_only) atomicAdd(p.stats + p.Cout + gc, stats_s[kMaxCout3 + gc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, bool RES, uint32_t FLAGS>
__global__ void __launch_bounds__(kC3Threads, 1) conv3_tc_kernel(const __grid_constant__ Conv3Params p) {
  conv3_tc_body<BN, RES, FLAGS, false>(p);
}

template <int BN, bool RES, uint32_t FLAGS>
__global__ void __launch_bounds__(kC3Threads, 1) conv3_tc_dual_kernel(const __grid_constant__ Conv3Params pa,
                                                                      const __grid_constant__ Conv3Params pb) {
  if (blockIdx.x & 1)
    conv3_tc_body<BN, RES, FLAGS, true>(pb);
  else
    conv3_tc_body<BN, RES, FLAGS, true>(pa);
}

}  // namespace eb
