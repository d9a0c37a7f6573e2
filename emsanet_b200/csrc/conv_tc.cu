// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (NHWC bf16, fp32 accumulate).
//
//   conv_tc_kernel  : forward conv and data-gradient (a dgrad is a conv with flipped taps and
//                     transposed weights) for the 1x1, 3x1, 1x3 and 3x3 filters of EMSANet
//                     (reference call sites: MT/model/block.py:174-190 NBt1D 3x1/1x3,
//                      MT/model/decoder/dense_base.py:43-46 3x3, MT/model/encoder_decoder_fusion.py:57-60 1x1,
//                      MT/model/context_module/ppm.py:43-54 1x1, MT/model/backbone/resnet.py:139-143 1x1 s2).
//   wgrad_tc_kernel : weight gradient, reduction over pixels on the tensor cores (MN-major operands).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> swizzled smem -> coalesced global stores).
// The CTA is persistent over output tiles; the accumulator is double-buffered in TMEM (2 x 256 columns)
// so the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
#include "conv_tc.cuh"
#include "ptx.cuh"
#include "conv3_tc.cuh"
#include "wgrad3_tc.cuh"
#include "conv3_2cta.cuh"

namespace eb {

constexpr int kThreads = 320;      // conv: 2 control warps + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kWgThreads = 192;    // wgrad: 2 control warps + 4 epilogue warps
constexpr int kATileBytes = 128 * 128;   // 128 pixels x 64 bf16
constexpr int kStageBufBytes = 8 * 1024;   // epilogue staging: 8 warps x 2 KB private slots = 2 x kStageBufBytes
constexpr int kMaxCout = 1024;             // per-CTA statistics scratch

// b_res_bytes > 0: weight-stationary mode — the stages hold only the A tile, the weights live in a separate region
__host__ __device__ inline int conv_stage_bytes(int block_n, int b_res_bytes = 0, int ksub = 1) {
  return ksub * (b_res_bytes > 0 ? kATileBytes : kATileBytes + block_n * 128);
}
__host__ __device__ inline int conv_smem_bytes(int block_n, int stages, int b_res_bytes = 0, int ksub = 1) {
  return stages * conv_stage_bytes(block_n, b_res_bytes, ksub) + b_res_bytes + 2 * kStageBufBytes +
         3 * kMaxCout * 4 + 256 + 1024;
}

// FLAGS: compile-time epilogue flags (kGenericFlags = read p.flags at run time).  The epilogue is instruction bound on
// the small-C layers, so every flag combination the network uses gets its own specialisation.
constexpr uint32_t kGenericFlags = 0xFFFFFFFFu;
#ifndef EB200_CONV_PROBES
#define EB200_CONV_PROBES 0   // 1: compile the clock64 trace / skip-switch hooks used by scripts/conv_probe.py
#endif
template <uint32_t FLAGS>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  const uint32_t flags = FLAGS == kGenericFlags ? p.flags : FLAGS;
  const int dbg = EB200_CONV_PROBES ? p.debug : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const int kblocks_ = p.Cin >> 6;
  const int b_res_bytes = p.b_resident ? p.taps * kblocks_ * p.block_n * 128 : 0;
  const int stage_bytes = conv_stage_bytes(p.block_n, b_res_bytes, p.ksub);
  const int sub_bytes = stage_bytes / p.ksub;
  const uint32_t pipe_base = smem_base;
  const uint32_t bres_base = pipe_base + p.stages * stage_bytes;  // resident weights (weight-stationary mode)
  const uint32_t stg_base = bres_base + b_res_bytes;              // 2 x 16 KB staging
  float* stats_s = reinterpret_cast<float*>(smem + p.stages * stage_bytes + b_res_bytes + 2 * kStageBufBytes);
  const float* bias_s = stats_s + 2 * kMaxCout;
  const uint32_t bar_base = stg_base + 2 * kStageBufBytes + 3 * kMaxCout * 4;
  // barriers: full[s] | empty[s] | tmem_full[2] | tmem_empty[2] | tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  const uint32_t bres_bar = bar_base + 8u * (2 * p.stages + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a[0]);
    tma_prefetch_desc(&p.map_a[1]);
    tma_prefetch_desc(&p.map_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), p.cluster);   // every CTA of the cluster releases the stage (its B slice lands in all)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiThreads);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
  }
  pdl_wait();   // set-up above overlaps the previous kernel's tail; global memory only from here on
  if (warp >= 2) {
    if (flags & kStats)
      for (int i = threadIdx.x - 64; i < 2 * kMaxCout; i += kEpiThreads) stats_s[i] = 0.f;
    if (flags & kBias)
      for (int i = threadIdx.x - 64; i < kMaxCout; i += kEpiThreads)
        stats_s[2 * kMaxCout + i] = i < p.Cout ? __ldg(p.bias + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();   // barrier inits visible before any peer multicasts / commits into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Work: "groups" of p.cluster consecutive pixel tiles with the same channel tile; cluster c takes groups
  // c, c + #clusters, ...; CTA rank r of the cluster takes pixel tile group*cluster + r (tiles past the end are
  // dummies: all loads out of bounds, nothing stored) so that every CTA of a cluster runs the same iterations.
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int cs = p.cluster;
  // CTA / cluster indices are read where each warp role starts its loop (opaque reads: hoisted to the top of the kernel
  // they end up in ordinary registers shared by the divergent roles and every uniform-datapath use pays an R2UR —
  // see conv3_tc.cuh)
#define EB_CONV_ROLE_INDICES                                                                     \
  const int crank = cs > 1 ? static_cast<int>(cluster_ctarank()) : 0;                            \
  const int group0 = cs > 1 ? static_cast<int>(cluster_id_x()) : c3_bx<false>();                 \
  const int gstride = cs > 1 ? static_cast<int>(cluster_count_x()) : c3_gd<false>();
  const int total_tiles = ((tiles_m + cs - 1) / cs) * p.tiles_c;   // number of groups
  const int kblocks = p.Cin >> 6;
  const int subs_per_tile = p.taps * kblocks;                       // 64-channel (tap, K-block) sub-blocks
  const int iters_per_tile = (subs_per_tile + p.ksub - 1) / p.ksub;  // pipeline stages per tile
  const uint16_t cmask = static_cast<uint16_t>((1u << cs) - 1u);
  const int bslice = p.block_n / cs;                                // weight rows fetched by each CTA

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop, lane 0 issues)
    uint32_t s = 0, ph = 0;
    if (p.b_resident && lane == 0) {   // all taps x channel blocks of the (single) channel tile, once per CTA
      mbar_arrive_expect_tx(bres_bar, b_res_bytes);
      for (int t = 0; t < p.taps; ++t)
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_3d(bres_base + (t * kblocks + kb) * p.block_n * 128, &p.map_b, bres_bar, kb * 64, 0, p.tap_w[t]);
    }
    EB_CONV_ROLE_INDICES
    for (int tile = group0; tile < total_tiles; tile += gstride) {
      const int mg = tile / p.tiles_c, ct = tile - mg * p.tiles_c;
      const int mt = mg * cs + crank;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int w0 = tw << p.lbw, h0 = th << p.lbh, n0 = tn << p.lbn;
      int t = 0, kb = 0;                 // (tap, channel block) of the next 64-channel sub-block
      for (int g = 0; g < iters_per_tile; ++g) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t sa = pipe_base + s * stage_bytes;
        const int nsub = min(p.ksub, subs_per_tile - g * p.ksub);
        if (dbg & 4) {
          if (lane == 0) mbar_arrive(full_bar(s));
        } else if (lane == 0) {
          mbar_arrive_expect_tx(full_bar(s), nsub * sub_bytes);
        }
        for (int j = 0; j < nsub; ++j) {
          if (lane == 0 && !(dbg & 4)) {
            tma_load_4d(sa + j * kATileBytes, &p.map_a[p.tap_view[t]], full_bar(s), kb * 64, w0 + p.tap_dx[t],
                        h0 + p.tap_dy[t], n0);
            if (p.b_resident) {
              // weights already in smem
            } else if (cs == 1) {
              tma_load_3d(sa + p.ksub * kATileBytes + j * p.block_n * 128, &p.map_b, full_bar(s), kb * 64,
                          ct * p.block_n, p.tap_w[t]);
            } else {   // my slice of the weight tile, delivered to every CTA of the cluster
              tma_load_3d_mc(sa + p.ksub * kATileBytes + j * p.block_n * 128 + crank * bslice * 128, &p.map_b,
                             full_bar(s), kb * 64, ct * p.block_n + crank * bslice, p.tap_w[t], cmask);
            }
          }
          if (++kb == kblocks) { kb = 0; ++t; }
        }
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, lane 0 issues)
    const uint32_t idesc = make_idesc_bf16(128, p.block_n, 0, 0);
    uint32_t s = 0, ph = 0, tl = 0;
    if (p.b_resident) {
      mbar_wait(bres_bar, 0);
      tc_fence_after();
    }
    EB_CONV_ROLE_INDICES
    for (int tile = group0; tile < total_tiles; tile += gstride, ++tl) {
      const uint32_t acc = tl & 1u, acc_ph = (tl >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;
      int i = 0;                         // running sub-block index within the tile
      for (int g = 0; g < iters_per_tile; ++g) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = pipe_base + s * stage_bytes;
        const int nsub = p.ksub == 1 ? 1 : min(p.ksub, subs_per_tile - g * p.ksub);
        for (int j = 0; j < nsub; ++j, ++i) {
          if (lane == 0) {
            const uint64_t adesc = make_smem_desc(sa + j * kATileBytes, 16, 1024);
            const uint64_t bdesc = make_smem_desc(
                p.b_resident ? bres_base + i * p.block_n * 128 : sa + p.ksub * kATileBytes + j * p.block_n * 128, 16,
                1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 4 x UMMA_K(16) = 64 channels; +32 B per step inside the 128 B swizzle row
              if (dbg & 2) break;
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
        }
        if (lane == 0) {
          if (cs == 1) umma_commit(empty_bar(s));
          else umma_commit_mc(empty_bar(s), cmask);
        }
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
      }
      if (lane == 0) umma_commit(tfull_bar(acc));
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 independent warps)
    // Warp (quarter, half) owns accumulator rows 32*quarter..+31 (its TMEM lane quarter) and columns
    // 32*half..+31 of every 64-column chunk.  It stages its 32x32 bf16 sub-tile (2 KB) in a PRIVATE smem slot to turn
    // "one row per lane" into 64-byte row segments per lane quad, so only __syncwarp is needed — no block barrier.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ewarp = warp - 2;
    const uint32_t wbuf = stg_base + ewarp * 2048;
    const bool has_bias = flags & kBias;
    const bool relu = flags & kRelu;
    const bool aux_add = flags & kAuxAdd;
    const bool aux_mask = flags & kAuxMask;
    const bool do_stats = flags & kStats;
    const bool relu_in_regs = relu && !aux_add;
    const int piece = lane & 3;                   // 16-byte piece (8 channels) of the 64-byte row segment
    const int srow = lane >> 2;                   // store rows srow + 8 j (within the warp's 32 rows)
    const bool stats_in_regs = do_stats && p.tiles_c == 1 && p.block_n <= 128;
    float rsum[2][8], rsq[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) rsum[i][k] = rsq[i][k] = 0.f;
    uint32_t tl = 0;
    EB_CONV_ROLE_INDICES
    for (int tile = group0; tile < total_tiles; tile += gstride, ++tl) {
      const int mg = tile / p.tiles_c, ct = tile - mg * p.tiles_c;
      const int mt = mg * cs + crank;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int w0 = tw << p.lbw, h0 = th << p.lbh, n0 = tn << p.lbn;
      const uint32_t acc = tl & 1u, acc_ph = (tl >> 1) & 1u;
      const int valid_cols = min(p.block_n, p.Cout - ct * p.block_n);   // multiple of 8
      const int nchunks = (p.block_n + 63) >> 6;

      long long ooff[4], aoff[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = quarter * 32 + srow + 8 * j;
        const int w = w0 + (r & ((1 << p.lbw) - 1));
        const int h = h0 + ((r >> p.lbw) & ((1 << p.lbh) - 1));
        const int n = n0 + (r >> (p.lbw + p.lbh));
        ok[j] = (w < p.W) && (h < p.H) && (n < p.N);
        ooff[j] = n * p.out_sn + h * p.out_sh + w * p.out_sw;
        if (aux_add || aux_mask) aoff[j] = n * p.aux_sn + h * p.aux_sh + w * p.aux_sw;
      }

      const bool etrace = (dbg & 8) && blockIdx.x == 0 && tl < 6 && warp == 2 && lane == 0;
      if (etrace) p.dbg_buf[tl * 64 + 32] = clock64();
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      if (etrace) p.dbg_buf[tl * 64 + 33] = clock64();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256u;

      for (int c = 0; c < nchunks; ++c) {
        if (etrace && c < 4) p.dbg_buf[tl * 64 + 34 + c] = clock64();
        const int colw = c * 64 + half * 32;                 // first column of this warp within the tile
        const int gcol = ct * p.block_n + colw + piece * 8;  // global channel of this lane's piece
        const bool pvalid = colw + piece * 8 < valid_cols;   // piece inside the real channels (warp-varying only by lane)
        const bool wactive = colw < p.block_n;               // this warp has columns in this chunk (warp-uniform)
        uint4 av[4];
        if ((aux_add || aux_mask) && wactive && pvalid) {    // issued first: latency hides behind the TMEM read
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ok[j]) av[j] = __ldg(reinterpret_cast<const uint4*>(p.aux + aoff[j] + gcol));
        }
        if (wactive) {
          uint32_t v[32];
          tmem_ld32(t_row + colw, v);
          tmem_ld_wait();
          if (c == nchunks - 1 || colw + 64 >= p.block_n + 32 * half) {
            // (release below, after the last TMEM read of this warp)
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (has_bias) {
            const float4* bp = reinterpret_cast<const float4*>(bias_s + ct * p.block_n + colw);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const float4 b4 = bp[q4];
              f[4 * q4 + 0] += b4.x; f[4 * q4 + 1] += b4.y; f[4 * q4 + 2] += b4.z; f[4 * q4 + 3] += b4.w;
            }
          }
          // lane == accumulator row: 4 x 16-byte pieces into the private slot, XOR-swizzled against bank conflicts
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
        if (c == nchunks - 1) {   // all TMEM reads of this accumulator by this thread are done
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
        }
        __syncwarp();
        if (wactive) {
          uint32_t x[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = srow + 8 * j;
            const uint32_t src = wbuf + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4);
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(x[j][0]), "=r"(x[j][1]), "=r"(x[j][2]), "=r"(x[j][3])
                         : "r"(src));
          }
          float ssum[8], ssq[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) ssum[k] = ssq[k] = 0.f;
          const bool bn_bwd = flags & kBnBwd;
          float bsc[8], bsh[8];
          if (bn_bwd && pvalid) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { bsc[k] = __ldg(p.bn_scale + gcol + k); bsh[k] = __ldg(p.bn_shift + gcol + k); }
          }
          if (pvalid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (!ok[j]) continue;
              if (aux_add || aux_mask) {
                const uint32_t a4[4] = {av[j].x, av[j].y, av[j].z, av[j].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float2 xv = unpack_bf16x2(x[j][k]);
                  const float2 a2 = unpack_bf16x2(a4[k]);
                  if (aux_add) {
                    xv.x += a2.x; xv.y += a2.y;
                    x[j][k] = relu ? pack_bf16x2_relu(xv.x, xv.y) : pack_bf16x2(xv.x, xv.y);
                  } else if (bn_bwd) {
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
                const uint32_t a4b[4] = {av[j].x, av[j].y, av[j].z, av[j].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 xv = unpack_bf16x2(x[j][k]);
                  float2 m2 = xv;                               // second moment partner: the value itself, or x (= aux)
                  if (bn_bwd) m2 = unpack_bf16x2(a4b[k]);
                  ssum[2 * k] += xv.x; ssq[2 * k] = fmaf(xv.x, m2.x, ssq[2 * k]);
                  ssum[2 * k + 1] += xv.y; ssq[2 * k + 1] = fmaf(xv.y, m2.y, ssq[2 * k + 1]);
                }
              }
              if (!(dbg & 1))
                *reinterpret_cast<uint4*>(p.out + ooff[j] + gcol) = make_uint4(x[j][0], x[j][1], x[j][2], x[j][3]);
            }
          }
          if (stats_in_regs) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (c == 0) { rsum[0][k] += ssum[k]; rsq[0][k] += ssq[k]; }
              else { rsum[1][k] += ssum[k]; rsq[1][k] += ssq[k]; }
            }
          } else if (do_stats) {
            // lanes with equal piece own the same 8 channels
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                ssum[k] += __shfl_xor_sync(0xffffffffu, ssum[k], off);
                ssq[k] += __shfl_xor_sync(0xffffffffu, ssq[k], off);
              }
            }
            if (lane < 4 && pvalid) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                atomicAdd(&stats_s[gcol + k], ssum[k]);
                atomicAdd(&stats_s[kMaxCout + gcol + k], ssq[k]);
              }
            }
          }
        }
        __syncwarp();   // the slot is rewritten by the next chunk
      }
      if (etrace) p.dbg_buf[tl * 64 + 40] = clock64();
    }
    if (stats_in_regs) {   // one reduction round for the whole CTA (tiles_c == 1: channels are tile-invariant)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            rsum[i][k] += __shfl_xor_sync(0xffffffffu, rsum[i][k], off);
            rsq[i][k] += __shfl_xor_sync(0xffffffffu, rsq[i][k], off);
          }
        }
        const int gc = i * 64 + half * 32 + piece * 8;
        if (lane < 4 && gc < p.Cout) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            atomicAdd(&stats_s[gc + k], rsum[i][k]);
            atomicAdd(&stats_s[kMaxCout + gc + k], rsq[i][k]);
          }
        }
      }
    }
    if (do_stats) {
      named_bar_sync(1, kEpiThreads);
      for (int cidx = threadIdx.x - 64; cidx < p.Cout; cidx += kEpiThreads) {
        atomicAdd(p.stats + cidx, stats_s[cidx]);
        if (!(flags & kStatsSum)) atomicAdd(p.stats + p.Cout + cidx, stats_s[kMaxCout + cidx]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();   // no CTA leaves while peers may still multicast into it / signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
#undef EB_CONV_ROLE_INDICES

// ------------------------------------------------------------------------------------------------
// Weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int kWgPix = 64;                        // pixels (GEMM K) per pipeline stage
constexpr int kWgSubTile = kWgPix * 128;          // 64 pixels x 64 channels bf16 = 8 KB

__host__ __device__ inline int wgrad_stage_bytes(int block_n, int taps) {
  return 2 * kWgSubTile + taps * (block_n / 64) * kWgSubTile;
}
__host__ __device__ inline int wgrad_smem_bytes(int block_n, int taps, int stages) {
  return stages * wgrad_stage_bytes(block_n, taps) + 256 + 1024;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stage_bytes = wgrad_stage_bytes(p.block_n, p.taps);
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // work item decode: blockIdx.x = ((co_tile * ci_tiles + ci_tile) * tap_groups + tg) * ksplit + ks
  const int ci_tiles = (p.Cin + p.block_n - 1) / p.block_n;
  int bid = blockIdx.x;
  const int ks = bid % p.ksplit; bid /= p.ksplit;
  const int tg = bid % p.tap_groups; bid /= p.tap_groups;
  const int ci_tile = bid % ci_tiles;
  const int co_tile = bid / ci_tiles;
  const int tap0 = tg * p.taps;
  const int ntaps = min(p.taps, p.total_taps - tap0);

  const int total_boxes = p.tiles_w * p.tiles_h * p.tiles_n;
  const int per = (total_boxes + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(total_boxes, kb0 + per);
  const int nk = max(0, kb1 - kb0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_dy);
    tma_prefetch_desc(&p.map_x[0]);
    tma_prefetch_desc(&p.map_x[1]);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nsub = p.block_n >> 6;   // 64-channel sub-tiles of the ci block

  if (nk > 0) {
    if (warp == 0) {
      // warp-uniform producer loop (lane 0 issues); box coordinates advance incrementally
      const uint32_t tx_bytes = 2 * kWgSubTile + ntaps * nsub * kWgSubTile;
      uint32_t s = 0, ph = 0;
      int tw = kb0 % p.tiles_w;
      int th = (kb0 / p.tiles_w) % p.tiles_h;
      int tn = kb0 / (p.tiles_w * p.tiles_h);
      for (int i = 0; i < nk; ++i) {
        const int w0 = tw << p.lbw, h0 = th << p.lbh, n0 = tn << p.lbn;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (lane == 0) {
          const uint32_t sa = smem_base + s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), tx_bytes);
          tma_load_4d(sa, &p.map_dy, full_bar(s), co_tile * 128, w0, h0, n0);
          tma_load_4d(sa + kWgSubTile, &p.map_dy, full_bar(s), co_tile * 128 + 64, w0, h0, n0);
          for (int t = 0; t < ntaps; ++t) {
            const int ta = tap0 + t;
            const CUtensorMap* mx = &p.map_x[p.tap_view[ta]];
            for (int j = 0; j < nsub; ++j) {
              tma_load_4d(sa + (2 + t * nsub + j) * kWgSubTile, mx, full_bar(s), ci_tile * p.block_n + j * 64,
                          w0 + p.tap_dx[ta], h0 + p.tap_dy[ta], n0);
            }
          }
        }
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
        if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_bf16(128, p.block_n, 1, 1);
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < nk; ++i) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + s * stage_bytes;
          for (int t = 0; t < ntaps; ++t) {
            const uint32_t sb = sa + (2 + t * nsub) * kWgSubTile;
#pragma unroll
            for (int k = 0; k < kWgPix / 16; ++k) {   // 16 pixels (2 swizzle atoms of 8 rows) per UMMA
              const uint64_t adesc = make_smem_desc(sa + k * 2048, kWgSubTile, 1024);
              const uint64_t bdesc = make_smem_desc(sb + k * 2048, kWgSubTile, 1024);
              umma_bf16(tmem_base + t * p.block_n, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(s));
        }
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
      }
      if (lane == 0) umma_commit(tfull_bar);
      __syncwarp();
    } else {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const int co = co_tile * 128 + row;
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      // Fast path (p.bulk): thread r owns accumulator row r = output channel co.  Its ntaps x block_n results are ONE
      // contiguous run of the destination — dW[co][ci0..][0..2] for 3-tap filters, dW[co][ci0..] for 1x1, and for the
      // 3 tap groups of a 3x3 filter a run of the staging workspace ws[tg][co][ci][3] (folded into dW[co][ci][9] by
      // wgrad_ws_finish_kernel) — so it is interleaved into the thread's own smem row (the pipeline buffers are dead
      // once tfull fires) and handed to the TMA as one bulk reduce-add: the L2 adds whole lines.
      if (p.bulk) {
        const int nt = p.bulk == 1 ? 1 : 3;
        const int ci0 = ci_tile * p.block_n;
        const int ncols = min(p.block_n, p.Cin - ci0);                 // valid ci of this tile (multiple of 4)
        const uint32_t pitch = static_cast<uint32_t>(nt * p.block_n * 4 + 16);
        const uint32_t srow = smem_base + row * pitch;
        for (int g = 0; g < (p.block_n >> 4); ++g) {
          if (nt == 3) {
            uint32_t v[3][16];
            tmem_ld16(t_row + 0 * p.block_n + g * 16, v[0]);
            tmem_ld16(t_row + 1 * p.block_n + g * 16, v[1]);
            tmem_ld16(t_row + 2 * p.block_n + g * 16, v[2]);
            tmem_ld_wait();
            uint32_t o[48];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              o[3 * j + 0] = v[0][j];
              o[3 * j + 1] = v[1][j];
              o[3 * j + 2] = v[2][j];
            }
#pragma unroll
            for (int q4 = 0; q4 < 12; ++q4)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(srow + g * 192 + q4 * 16), "r"(o[4 * q4]),
                           "r"(o[4 * q4 + 1]), "r"(o[4 * q4 + 2]), "r"(o[4 * q4 + 3]) : "memory");
          } else {
            uint32_t v[16];
            tmem_ld16(t_row + g * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(srow + g * 64 + q4 * 16), "r"(v[4 * q4]),
                           "r"(v[4 * q4 + 1]), "r"(v[4 * q4 + 2]), "r"(v[4 * q4 + 3]) : "memory");
          }
        }
        fence_proxy_async();
        if (co < p.Cout && ncols > 0 && p.dw != nullptr) {
          float* dst;
          if (p.bulk == 3) dst = p.ws + ((static_cast<long long>(tg) * p.Cout + co) * p.Cin + ci0) * 3;
          else dst = p.dw + co * p.dw_sco + static_cast<long long>(ci0) * nt;
          bulk_reduce_add_f32(dst, srow, static_cast<uint32_t>(nt * ncols * 4));
          bulk_commit_group();
          bulk_wait_group_read0();
        }
      } else {
      for (int t = 0; t < ntaps; ++t) {
        for (int g = 0; g < (p.block_n >> 4); ++g) {
          uint32_t v[16];
          tmem_ld16(t_row + t * p.block_n + g * 16, v);
          tmem_ld_wait();
          if (co < p.Cout && p.dw != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ci = ci_tile * p.block_n + g * 16 + j;
              if (ci < p.Cin) {
                atomicAdd(p.dw + co * p.dw_sco + ci * p.dw_sci + (tap0 + t) * p.dw_st, __uint_as_float(v[j]));
              }
            }
          }
        }
      }
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

}  // namespace eb

// ================================================================================================
// Host side: tensor maps, tiling heuristics, C-ABI entry points
// ================================================================================================
#include "../../include/emsanet_b200.h"
#include "common.h"
#include <cstdlib>
#include <cstring>

namespace eb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// 4-D map over an NHWC bf16 view: dims (C, W, H, N), box (64, bw, bh, bn), 128-byte swizzle, zero OOB fill.
static int make_view_map(CUtensorMap* m, const eb200_view& v, int bw, int bh, int bn) {
  EncodeTiledFn fn = encode_fn();
  EB_REQUIRE(fn, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  EB_REQUIRE(v.ptr && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, "view pointer must be 16-byte aligned");
  EB_REQUIRE((v.sw * 2) % 16 == 0 && (v.sh * 2) % 16 == 0 && (v.sn * 2) % 16 == 0,
             "view strides must be multiples of 16 bytes (c=%d sw=%lld)", v.c, v.sw);
  cuuint64_t dims[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(view c=%d w=%d h=%d n=%d) failed: %d", v.c, v.w, v.h, v.n,
             (int)r);
  return 0;
}

static int make_weight_map(CUtensorMap* m, const void* w, int cin_pad, int cout_pad, int taps, int block_n) {
  EncodeTiledFn fn = encode_fn();
  EB_REQUIRE(fn, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * cout_pad * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)block_n, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight %dx%dx%d) failed: %d", cin_pad, cout_pad, taps, (int)r);
  return 0;
}

// pixel box (bw, bh, bn), bw*bh*bn == 2^lg, minimising padded work; ties -> wider rows
static void choose_box(int W, int H, int N, int lg, int* lbw, int* lbh, int* lbn) {
  double best = 1e30;
  for (int a = lg; a >= 0; --a) {
    for (int b = lg - a; b >= 0; --b) {
      const int c = lg - a - b;
      const double padded = (double)ceil_div(W, 1 << a) * (1 << a) * ceil_div(H, 1 << b) * (1 << b) *
                            ceil_div(N, 1 << c) * (1 << c);
      if (padded < best * 0.999) {
        best = padded;
        *lbw = a; *lbh = b; *lbn = c;
      }
    }
  }
}

static int g_smem_optin = 0;
static int smem_limit() {
  if (!g_smem_optin) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (g_smem_optin <= 0) g_smem_optin = 227 * 1024;
  }
  return g_smem_optin;
}

}  // namespace eb


// ------------------------------------------------------------------------------------------------
// 3-tap 1-D stride-1 convolutions (and their data gradients): halo kernel of conv3_tc.cuh
// ------------------------------------------------------------------------------------------------
namespace eb {

template <int BN, bool RES>
static void* conv3_kernel_for(uint32_t flags) {
  switch (flags) {
    case 0: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, 0>);
    case kBias: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, kBias>);
    case kBias | kRelu: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, kBias | kRelu>);
    case kAuxAdd: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, kAuxAdd>);
    case kBias | kAuxAdd | kRelu:   // inference with BatchNorm folded into the weights: conv + shift + residual + ReLU
      return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, kBias | kAuxAdd | kRelu>);
    case kStats: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, kStats>);
    case kAuxMask | kStats | kStatsSum: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, kAuxMask | kStats | kStatsSum>);
    case kAuxMask | kStats | kBnBwd: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, kAuxMask | kStats | kBnBwd>);
    default: return reinterpret_cast<void*>(conv3_tc_kernel<BN, RES, 0xFFFFFFFFu>);
  }
}

static void* conv3_2cta_kernel_for(uint32_t flags) {
  switch (flags) {
    case 0: return reinterpret_cast<void*>(conv3_2cta_kernel<0>);
    case kBias: return reinterpret_cast<void*>(conv3_2cta_kernel<kBias>);
    case kBias | kRelu: return reinterpret_cast<void*>(conv3_2cta_kernel<kBias | kRelu>);
    case kAuxAdd: return reinterpret_cast<void*>(conv3_2cta_kernel<kAuxAdd>);
    case kStats: return reinterpret_cast<void*>(conv3_2cta_kernel<kStats>);
    case kAuxMask | kStats | kStatsSum: return reinterpret_cast<void*>(conv3_2cta_kernel<kAuxMask | kStats | kStatsSum>);
    default: return reinterpret_cast<void*>(conv3_2cta_kernel<0xFFFFFFFFu>);
  }
}

template <uint32_t F>
static void* conv3_dual_fn() { return reinterpret_cast<void*>(conv3_tc_dual_kernel<256, false, F>); }
static void* conv3_dual_kernel_for(uint32_t flags) {
  switch (flags) {
    case 0: return conv3_dual_fn<0>();
    case kBias: return conv3_dual_fn<kBias>();
    case kBias | kRelu: return conv3_dual_fn<kBias | kRelu>();
    case kAuxAdd: return conv3_dual_fn<kAuxAdd>();
    case kBias | kAuxAdd | kRelu: return conv3_dual_fn<kBias | kAuxAdd | kRelu>();
    case kStats: return conv3_dual_fn<kStats>();
    case kAuxMask | kStats | kStatsSum: return conv3_dual_fn<kAuxMask | kStats | kStatsSum>();
    case kAuxMask | kStats | kBnBwd: return conv3_dual_fn<kAuxMask | kStats | kBnBwd>();
    default: return conv3_dual_fn<0xFFFFFFFFu>();
  }
}

struct Conv3Plan {
  Conv3Params p;
  void* fn;
  int BN, grid, npairs, smem;
  bool res, pair;
};

static int conv3_configure(void* fn) {
  static void* configured[96] = {};
  int i = 0;
  for (; i < 96 && configured[i] && configured[i] != fn; ++i) {}
  if (i < 96 && !configured[i]) {
    EB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()));
    configured[i] = fn;
  }
  return 0;
}

// Plans a launch of the halo kernel on `sms` SMs; *ok = false -> not eligible (use the generic kernel).
static int plan_conv3(const eb200_conv_desc* d, int sms, Conv3Plan* plan, bool* handled) {
  *handled = false;
  if (d->taps != 3 || getenv("EB200_CONV3_DISABLE")) return 0;   // (env read per call: the tests toggle it)
  bool along_h = true, along_w = true;
  int seen = 0;
  for (int t = 0; t < 3; ++t) {
    if (d->tap_view[t] != 0) return 0;
    if (d->tap_dx[t] != 0) along_h = false;
    if (d->tap_dy[t] != 0) along_w = false;
    const int o = d->tap_dx[t] + d->tap_dy[t];
    if (o < -1 || o > 1) return 0;
    seen |= 1 << (o + 1);
  }
  if (along_h == along_w || seen != 7) return 0;
  if (d->cout != d->cout_pad || d->cout % 64 != 0 || d->cin_pad % 64 != 0 || d->cout > kMaxCout3) return 0;
  int BN = d->cout >= 256 ? 256 : d->cout;
  if (const char* e = getenv("EB200_CONV3_BN")) {          // experiments: force the channel tile of the wide layers
    const int f = atoi(e);
    if ((f == 128 || f == 256) && d->cout >= 256 && d->cout % f == 0) BN = f;
  }
  if (BN != 64 && BN != 128 && BN != 256) return 0;
  if (d->cout % BN != 0) return 0;
  const int ext_f = along_h ? d->w : d->h;     // fast axis = the one the taps do NOT move along
  const int ext_s = along_h ? d->h : d->w;
  if (ext_f < 8) return 0;
  const long long max_off = (long long)d->n * d->out_sn;
  const long long max_aoff = (d->flags & (EB200_AUX_ADD | EB200_AUX_MASK)) ? (long long)d->n * d->aux_sn : 0;
  if (max_off >= (1ll << 31) || max_aoff >= (1ll << 31)) return 0;   // 32-bit element offsets in the epilogue

  Conv3Params& p = plan->p;
  memset(&p, 0, sizeof(p));
  // tile shape: fewest tiles, then least halo overhead (largest S)
  long long best = -1;
  for (int lg = 3; lg <= 5; ++lg) {
    const int F = 1 << lg, S = 128 >> lg;
    const long long tiles = (long long)ceil_div(ext_f, F) * ceil_div(ext_s, S);
    if (best < 0 || tiles < best) { best = tiles; p.lgF = lg; }
  }
  const int F = 1 << p.lgF, S = 128 >> p.lgF;
  p.N = d->n; p.ext_f = ext_f; p.ext_s = ext_s;
  p.Cout = d->cout; p.kblocks = d->cin_pad / 64;
  p.tiles_f = ceil_div(ext_f, F); p.tiles_s = ceil_div(ext_s, S); p.tiles_c = d->cout / BN;
  const long long tiles_m = (long long)p.tiles_f * p.tiles_s * d->n;
  if (tiles_m * p.tiles_c >= (1ll << 30)) return 0;
  p.total_tiles = static_cast<int>(tiles_m * p.tiles_c);
  for (int t = 0; t < 3; ++t) {
    p.tap_row[t] = (d->tap_dx[t] + d->tap_dy[t] + 1) * F;
    p.tap_w[t] = d->tap_w[t];
    EB_REQUIRE(d->tap_w[t] >= 0 && d->tap_w[t] < d->weight_taps, "eb200_conv2d: tap_w[%d]=%d out of %d", t, d->tap_w[t], d->weight_taps);
  }
  p.a_bytes = (S + 2) * F * 128;
  const int w_bytes = 3 * p.kblocks * BN * 128;
  const int budget = smem_limit() - conv3_fixed_smem();
  int grid = p.total_tiles < sms ? p.total_tiles : sms;
  grid -= grid % p.tiles_c;                    // a CTA keeps its channel tile (register-resident statistics)
  if (grid < p.tiles_c) return 0;
  bool res = BN <= 128 && p.tiles_c == 1 && w_bytes <= 100 * 1024 && p.total_tiles >= 2 * grid &&
             budget - w_bytes >= 3 * p.a_bytes;
  if (BN == 64 && !res) return 0;              // tiny C=64 problems stay on the generic kernel
  if (BN == 256) res = false;
  // wide layers: CTA pairs (cta_group::2) share the weight tile, M = 256 pixels per pair
  // (only with >= 4 pair tiles per pair, e.g. the 1024x768 configuration: measured +3..8 % there, but at config-2 sizes
  // — 2 tiles per SM — the cluster launch and the cross-CTA barrier latency cost as much as the halved weight traffic
  // saves (scripts/microbench_2cta.py, scripts/ab_pdl.sh); EB200_CONV3_2CTA=1 / EB200_CONV3_NO_2CTA=1 force either)
  const bool pair = BN == 256 && sms == num_sms() && num_sms() >= 2 * p.tiles_c && !getenv("EB200_CONV3_NO_2CTA") &&
                    (getenv("EB200_CONV3_2CTA") || ((tiles_m + 1) / 2) * p.tiles_c >= 2 * num_sms());
  int npairs = 0;
  if (pair) {
    const long long items = ((tiles_m + 1) / 2) * p.tiles_c;
    p.total_tiles = static_cast<int>(items);
    npairs = items < num_sms() / 2 ? static_cast<int>(items) : num_sms() / 2;
    npairs -= npairs % p.tiles_c;
  }
  if (pair) {
    p.stages_b = 6;                            // 6 x (128 rows x 128 B) per CTA
    p.stages_a = (budget - p.stages_b * 128 * 128) / p.a_bytes;
  } else if (res) {
    p.stages_a = (budget - w_bytes) / p.a_bytes;
    p.stages_b = 1;
  } else {
    p.stages_b = BN == 256 ? 4 : 6;
    p.stages_a = (budget - p.stages_b * BN * 128) / p.a_bytes;
  }
  if (p.stages_a > kC3MaxStages) p.stages_a = kC3MaxStages;
  if (p.stages_a < 2) return 0;
  p.flags = d->flags;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_sn = d->out_sn; p.out_ss = along_h ? d->out_sh : d->out_sw; p.out_sf = along_h ? d->out_sw : d->out_sh;
  p.aux = static_cast<const __nv_bfloat16*>(d->aux);
  p.aux_sn = d->aux_sn; p.aux_ss = along_h ? d->aux_sh : d->aux_sw; p.aux_sf = along_h ? d->aux_sw : d->aux_sh;
  p.bias = d->bias;
  p.stats = d->stats;
  p.bn_scale = d->bn_scale;
  p.bn_shift = d->bn_shift;

  eb200_view v = d->in[0];                     // dims (C, fast, slow, N): the W<->H permuted view for 1x3 filters
  if (!along_h) {
    v.w = d->in[0].h; v.h = d->in[0].w; v.sw = d->in[0].sh; v.sh = d->in[0].sw;
  }
  if (make_view_map(&p.map_a, v, F, S + 2, 1)) return 1;
  if (make_weight_map(&p.map_b, d->weight, d->cin_pad, d->cout_pad, d->weight_taps, pair ? BN / 2 : BN)) return 1;

  void* fn = nullptr;
  if (pair) fn = conv3_2cta_kernel_for(d->flags);
  else if (BN == 64) fn = conv3_kernel_for<64, true>(d->flags);
  else if (BN == 128) fn = res ? conv3_kernel_for<128, true>(d->flags) : conv3_kernel_for<128, false>(d->flags);
  else fn = conv3_kernel_for<256, false>(d->flags);
  plan->fn = fn;
  plan->BN = BN; plan->grid = grid; plan->npairs = npairs; plan->res = res; plan->pair = pair;
  plan->smem = p.stages_a * p.a_bytes +
               (pair ? p.stages_b * 128 * 128 : res ? w_bytes : p.stages_b * BN * 128) + conv3_fixed_smem();
  *handled = true;
  return 0;
}

// returns 0 and sets *handled when the launch was made by the halo kernel; *handled = false -> use the generic kernel
static int launch_conv3(const eb200_conv_desc* d, void* stream, bool* handled) {
  Conv3Plan plan;
  if (plan_conv3(d, num_sms(), &plan, handled)) return 1;
  if (!*handled) return 0;
  if (conv3_configure(plan.fn)) return 1;
  if (plan.pair) {
    void* args[1] = {&plan.p};
    EB_CUDA(launch_ex(plan.fn, dim3(2 * plan.npairs), dim3(kC3Threads), plan.smem, static_cast<cudaStream_t>(stream), args, 2));
    return launch_check("conv3_2cta_kernel");
  }
  void* args[1] = {&plan.p};
  EB_CUDA(launch_ex(plan.fn, dim3(plan.grid), dim3(kC3Threads), plan.smem, static_cast<cudaStream_t>(stream), args));
  return launch_check("conv3_tc_kernel");
}

// Two independent convolutions of identical geometry in ONE launch (even / odd CTAs); *handled = false -> the caller
// launches them one after the other.
static int launch_conv3_dual(const eb200_conv_desc* a, const eb200_conv_desc* b, void* stream, bool* handled) {
  *handled = false;
  if (getenv("EB200_NO_DUAL")) return 0;
  if (a->n != b->n || a->h != b->h || a->w != b->w || a->cin_pad != b->cin_pad || a->cout != b->cout ||
      a->flags != b->flags || a->taps != b->taps)
    return 0;
  for (int t = 0; t < a->taps && t < EB200_MAX_TAPS; ++t)
    if (a->tap_dx[t] != b->tap_dx[t] || a->tap_dy[t] != b->tap_dy[t] || a->tap_view[t] != b->tap_view[t]) return 0;
  Conv3Plan pa, pb;
  bool oka = false, okb = false;
  if (plan_conv3(a, num_sms() / 2, &pa, &oka)) return 1;
  if (!oka) return 0;
  if (plan_conv3(b, num_sms() / 2, &pb, &okb)) return 1;
  if (!okb || pa.BN != 256 || pb.BN != 256 || pa.res || pb.res || pa.pair || pb.pair || pa.grid != pb.grid ||
      pa.smem != pb.smem || pa.p.stages_a != pb.p.stages_a || pa.p.stages_b != pb.p.stages_b)
    return 0;
  void* fn = conv3_dual_kernel_for(a->flags);
  if (conv3_configure(fn)) return 1;
  void* args[2] = {&pa.p, &pb.p};
  EB_CUDA(launch_ex(fn, dim3(2 * pa.grid), dim3(kC3Threads), pa.smem, static_cast<cudaStream_t>(stream), args));
  *handled = true;
  return launch_check("conv3_tc_kernel(dual)");
}

}  // namespace eb


// ------------------------------------------------------------------------------------------------
// weight gradient of the 3-tap 1-D stride-1 convolutions: halo kernel of wgrad3_tc.cuh
// ------------------------------------------------------------------------------------------------
namespace eb {

struct Wgrad3Plan {
  Wgrad3Params p;
  int BN, grid, smem, cluster;
};

static int plan_wgrad3(const eb200_wgrad_desc* d, int sms, Wgrad3Plan* plan, bool* handled) {
  *handled = false;
  if (d->taps != 3 || getenv("EB200_WGRAD3_DISABLE")) return 0;
  bool along_h = true, along_w = true;
  for (int t = 0; t < 3; ++t) {
    if (d->tap_view[t] != 0) return 0;
    if (d->tap_dx[t] != 0) along_h = false;
    if (d->tap_dy[t] != 0) along_w = false;
    if (d->tap_dx[t] + d->tap_dy[t] != t - 1) return 0;    // TMEM group t == absolute tap t (offsets -1, 0, +1)
  }
  if (along_h == along_w) return 0;
  const int cin = d->x[0].c, cout = d->dy.c;
  if (cin % 64 != 0 || cout % 8 != 0) return 0;
  if (d->dw_st != 1 || d->dw_sci != 3 || (d->dw_sco & 3) != 0 || (reinterpret_cast<uintptr_t>(d->dw) & 15) != 0) return 0;
  const int BN = cin >= 128 ? 128 : 64;
  if (cin % BN != 0) return 0;
  const int ext_f = along_h ? d->dy.w : d->dy.h;
  const int ext_s = along_h ? d->dy.h : d->dy.w;
  if (ext_f < 8) return 0;

  Wgrad3Params& p = plan->p;
  memset(&p, 0, sizeof(p));
  long long best = -1;
  for (int lg = 3; lg <= 5; ++lg) {
    const int F = 1 << lg, S = 64 >> lg;
    const long long boxes = (long long)ceil_div(ext_f, F) * ceil_div(ext_s, S);
    if (best < 0 || boxes < best) { best = boxes; p.lgF = lg; }
  }
  const int F = 1 << p.lgF, S = 64 >> p.lgF;
  p.N = d->dy.n; p.ext_f = ext_f; p.ext_s = ext_s;
  p.Cin = cin; p.Cout = cout;
  p.tiles_f = ceil_div(ext_f, F); p.tiles_s = ceil_div(ext_s, S);
  const long long boxes = (long long)p.tiles_f * p.tiles_s * p.N;
  if (boxes >= (1ll << 30)) return 0;
  p.total_boxes = static_cast<int>(boxes);
  p.ci_tiles = cin / BN;
  const int items = ceil_div(cout, 128) * p.ci_tiles;
  int ksplit = (sms + items / 2) / items;
  if (ksplit > ceil_div(p.total_boxes, 4)) ksplit = ceil_div(p.total_boxes, 4);
  if (ksplit < 1) ksplit = 1;
  while (ksplit > 1 && ceil_div(p.total_boxes, ksplit) * (ksplit - 1) >= p.total_boxes) --ksplit;
  if (const char* e = getenv("EB200_WGRAD_KSPLIT")) { const int k = atoi(e); if (k >= 1 && k <= p.total_boxes) ksplit = k; }
  // CTA pairs (cluster of 2) hold adjacent K-splits of one output tile and pre-reduce through DSMEM (needs an even split).
  // Opt-in (EB200_WGRAD_CLUSTER=1): measured +1.5 ms per step at config-2 sizes — the cluster launch and the two
  // cluster barriers cost more than the halved L2 reduction traffic saves (scripts/ab_env.sh).
  plan->cluster = 1;
  if (sms == num_sms() && ksplit >= 4 && getenv("EB200_WGRAD_CLUSTER")) {
    ksplit -= ksplit & 1;
    plan->cluster = 2;
  }
  p.ksplit = ksplit;
  for (int t = 0; t < 3; ++t) p.tap_row[t] = t * F;
  p.x_sub_bytes = (S + 2) * F * 128;
  const int stage_bytes = 2 * kWg3DySub + (BN / 64) * p.x_sub_bytes;
  int stages = (smem_limit() - 1024 - 256) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2 || stages * stage_bytes < 128 * (3 * BN * 4 + 16)) return 0;   // the epilogue re-uses the stage buffers
  p.stages = stages;
  p.dw = getenv("EB200_WGRAD_NOSTORE") ? nullptr : d->dw;   // experiments: main loop without the reductions
  p.dw_sco = d->dw_sco;

  eb200_view vdy = d->dy, vx = d->x[0];
  if (!along_h) {
    vdy.w = d->dy.h; vdy.h = d->dy.w; vdy.sw = d->dy.sh; vdy.sh = d->dy.sw;
    vx.w = d->x[0].h; vx.h = d->x[0].w; vx.sw = d->x[0].sh; vx.sh = d->x[0].sw;
  }
  if (make_view_map(&p.map_dy, vdy, F, S, 1)) return 1;
  if (make_view_map(&p.map_x, vx, F, S + 2, 1)) return 1;
  plan->BN = BN;
  plan->grid = items * ksplit;
  plan->smem = stages * stage_bytes + 1024 + 256;
  *handled = true;
  return 0;
}

static int wgrad3_configure(void* fn) {
  static void* configured[8] = {};
  int i = 0;
  for (; i < 8 && configured[i] && configured[i] != fn; ++i) {}
  if (i < 8 && !configured[i]) {
    EB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()));
    configured[i] = fn;
  }
  return 0;
}

static int launch_wgrad3(const eb200_wgrad_desc* d, void* stream, bool* handled) {
  Wgrad3Plan plan;
  if (plan_wgrad3(d, num_sms(), &plan, handled)) return 1;
  if (!*handled) return 0;
  void* fn;
  if (plan.cluster == 2)
    fn = plan.BN == 128 ? reinterpret_cast<void*>(wgrad3_tc_kernel<128, 2>)
                        : reinterpret_cast<void*>(wgrad3_tc_kernel<64, 2>);
  else
    fn = plan.BN == 128 ? reinterpret_cast<void*>(wgrad3_tc_kernel<128>) : reinterpret_cast<void*>(wgrad3_tc_kernel<64>);
  if (wgrad3_configure(fn)) return 1;
  void* args[1] = {&plan.p};
  EB_CUDA(launch_ex(fn, dim3(plan.grid), dim3(kWg3Threads), plan.smem, static_cast<cudaStream_t>(stream), args,
                    plan.cluster, 2));
  return launch_check("wgrad3_tc_kernel");
}

static int launch_wgrad3_dual(const eb200_wgrad_desc* a, const eb200_wgrad_desc* b, void* stream, bool* handled) {
  *handled = false;
  // Off by default: split-K already fills all SMs with ONE weight gradient (no wave quantisation to recover), so two
  // problems on half the SMs each only lengthen the main loop — measured 61 us per double launch against 2 x 25 us
  // (profiles/r1_launches_step_v6.csv).  EB200_WGRAD_DUAL=1 enables it (the parity test does).
  if (getenv("EB200_NO_DUAL") || !getenv("EB200_WGRAD_DUAL")) return 0;
  if (a->dy.n != b->dy.n || a->dy.h != b->dy.h || a->dy.w != b->dy.w || a->dy.c != b->dy.c || a->x[0].c != b->x[0].c ||
      a->taps != b->taps)
    return 0;
  for (int t = 0; t < a->taps && t < EB200_MAX_TAPS; ++t)
    if (a->tap_dx[t] != b->tap_dx[t] || a->tap_dy[t] != b->tap_dy[t] || a->tap_view[t] != b->tap_view[t]) return 0;
  Wgrad3Plan pa, pb;
  bool oka = false, okb = false;
  if (plan_wgrad3(a, num_sms() / 2, &pa, &oka)) return 1;
  if (!oka) return 0;
  if (plan_wgrad3(b, num_sms() / 2, &pb, &okb)) return 1;
  if (!okb || pa.BN != 128 || pb.BN != 128 || pa.grid != pb.grid || pa.smem != pb.smem || pa.p.stages != pb.p.stages)
    return 0;
  void* fn = reinterpret_cast<void*>(wgrad3_tc_dual_kernel<128>);
  if (wgrad3_configure(fn)) return 1;
  void* args[2] = {&pa.p, &pb.p};
  EB_CUDA(launch_ex(fn, dim3(2 * pa.grid), dim3(kWg3Threads), pa.smem, static_cast<cudaStream_t>(stream), args, 1, 2));
  *handled = true;
  return launch_check("wgrad3_tc_kernel(dual)");
}

}  // namespace eb

using namespace eb;

extern "C" int eb200_conv2d(const eb200_conv_desc* d, void* stream) {
  EB_REQUIRE(d && d->out && d->weight && d->in[0].ptr, "eb200_conv2d: null argument");
  EB_REQUIRE(d->taps >= 1 && d->taps <= kMaxTaps, "eb200_conv2d: taps=%d", d->taps);
  EB_REQUIRE(d->cout % 8 == 0 && d->cout <= kMaxCout, "eb200_conv2d: cout=%d must be a multiple of 8 and <= 1024", d->cout);
  EB_REQUIRE(d->cin_pad % 64 == 0 && d->cin_pad >= d->cin, "eb200_conv2d: cin_pad=%d", d->cin_pad);
  EB_REQUIRE(d->cout_pad % 16 == 0 && d->cout_pad >= d->cout, "eb200_conv2d: cout_pad=%d", d->cout_pad);
  EB_REQUIRE(!(d->flags & EB200_BIAS) || d->bias, "eb200_conv2d: bias flag without pointer");
  EB_REQUIRE(!(d->flags & (EB200_AUX_ADD | EB200_AUX_MASK)) || d->aux, "eb200_conv2d: aux flag without pointer");
  EB_REQUIRE(!(d->flags & EB200_STATS) || d->stats, "eb200_conv2d: stats flag without pointer");
  EB_REQUIRE(!(d->flags & EB200_BN_BWD) || ((d->flags & EB200_AUX_MASK) && (d->flags & EB200_STATS) && d->bn_scale &&
                                             d->bn_shift && !(d->flags & EB200_STATS_SUM_ONLY)),
             "eb200_conv2d: EB200_BN_BWD needs EB200_AUX_MASK | EB200_STATS and the BatchNorm affine");
  EB_REQUIRE((d->out_sw % 8) == 0 && (d->out_sh % 8) == 0 && (d->out_sn % 8) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
             "eb200_conv2d: output must be 16-byte aligned per pixel");
  {
    bool handled = false;
    if (launch_conv3(d, stream, &handled)) return 1;
    if (handled) return 0;
  }

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n; p.H = d->h; p.W = d->w;
  p.Cin = d->cin_pad; p.Cout = d->cout;
  p.taps = d->taps;
  bool use_view1 = false;
  for (int t = 0; t < d->taps; ++t) {
    p.tap_view[t] = d->tap_view[t]; p.tap_dy[t] = d->tap_dy[t]; p.tap_dx[t] = d->tap_dx[t]; p.tap_w[t] = d->tap_w[t];
    EB_REQUIRE(d->tap_view[t] == 0 || d->tap_view[t] == 1, "eb200_conv2d: tap_view");
    EB_REQUIRE(d->tap_w[t] >= 0 && d->tap_w[t] < d->weight_taps, "eb200_conv2d: tap_w[%d]=%d out of %d", t, d->tap_w[t], d->weight_taps);
    use_view1 |= d->tap_view[t] == 1;
  }
  choose_box(d->w, d->h, d->n, 7, &p.lbw, &p.lbh, &p.lbn);
  p.tiles_w = ceil_div(d->w, 1 << p.lbw);
  p.tiles_h = ceil_div(d->h, 1 << p.lbh);
  p.tiles_n = ceil_div(d->n, 1 << p.lbn);
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  // N tile: largest of 256/128/64 dividing cout_pad that minimises waves * (tile cost)
  int block_n = d->cout_pad;
  if (d->cout_pad > 64) {
    double best = 1e30;
    const int cands[3] = {256, 128, 64};
    for (int c : cands) {
      if (c > d->cout_pad || d->cout_pad % c) continue;
      // operand bytes streamed per CTA (A 128 rows + B c rows per K block) x rounds of tiles over the SMs
      const int tiles = tiles_m * (d->cout_pad / c);
      const double cost = (double)ceil_div(tiles, num_sms()) * (128 + c) * (c == 64 ? 1.15 : 1.0);
      if (cost < best * 0.999) { best = cost; block_n = c; }
    }
    if (best > 1e29) block_n = d->cout_pad <= 256 ? d->cout_pad : 0;
    static int force_bn = -1;
    if (force_bn < 0) { const char* e = getenv("EB200_CONV_BLOCKN"); force_bn = e ? atoi(e) : 0; }
    if (force_bn > 0 && d->cout_pad % force_bn == 0) block_n = force_bn;
  }
  EB_REQUIRE(block_n >= 16 && block_n <= 256 && block_n % 16 == 0 && d->cout_pad % block_n == 0,
             "eb200_conv2d: no N tile for cout_pad=%d", d->cout_pad);
  EB_REQUIRE(!(d->flags & EB200_STATS) || ((block_n % 64 == 0 || block_n == 32 || block_n == 16 || block_n == 96) &&
                                            d->cout % 16 == 0),
             "eb200_conv2d: stats need cout in 16-channel groups (cout=%d)", d->cout);
  p.block_n = block_n;
  p.tiles_c = ceil_div(d->cout, block_n);
  // weight-stationary mode: one channel tile whose complete weights fit next to >= 4 A stages
  int b_res_bytes = 0;
  {
    static int disable = -1;
    if (disable < 0) disable = getenv("EB200_CONV_NO_RESIDENT") ? 1 : 0;
    const int wbytes = d->taps * (d->cin_pad / 64) * block_n * 128;
    if (!disable && p.tiles_c == 1 && wbytes <= 112 * 1024 && tiles_m >= 2 * num_sms()) b_res_bytes = wbytes;
  }
  p.b_resident = b_res_bytes > 0;
  // sub-blocks per stage: as many as leave >= 2 stages (the MMA-issuing thread pays ~400 cycles per stage for the
  // barrier wait + tcgen05.commit and ~60 per MMA: narrow tiles are issue bound unless that is amortised)
  int ksub = 1;
  {
    static int fk = -1;
    if (fk < 0) { const char* e = getenv("EB200_CONV_KSUB"); fk = e ? atoi(e) : 0; }
    const int total_sub = d->taps * (d->cin_pad / 64);
    for (int k = 4; k >= 1; --k) {
      if (k > total_sub) continue;
      const int st = (smem_limit() - conv_smem_bytes(block_n, 0, b_res_bytes, k)) / conv_stage_bytes(block_n, b_res_bytes, k);
      if (st >= 2 || k == 1) { ksub = k; break; }   // measured: fewer, fatter stages win even at depth 2
    }
    if (fk > 0 && fk <= total_sub) ksub = fk;
  }
  p.ksub = ksub;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("EB200_CONV_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  const int fixed = conv_smem_bytes(block_n, 0, b_res_bytes, ksub);
  int stages = (smem_limit() - fixed) / conv_stage_bytes(block_n, b_res_bytes, ksub);
  if (stages > 8) stages = 8;
  { static int fs = -1; if (fs < 0) { const char* e = getenv("EB200_CONV_STAGES"); fs = e ? atoi(e) : 0; } if (fs > 0 && fs < stages) stages = fs; }
  EB_REQUIRE(stages >= 2, "eb200_conv2d: not enough shared memory");
  p.stages = stages;
  p.flags = d->flags;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_sn = d->out_sn; p.out_sh = d->out_sh; p.out_sw = d->out_sw;
  p.aux = static_cast<const __nv_bfloat16*>(d->aux);
  p.aux_sn = d->aux_sn; p.aux_sh = d->aux_sh; p.aux_sw = d->aux_sw;
  p.bias = d->bias;
  p.stats = d->stats;
  p.bn_scale = d->bn_scale;
  p.bn_shift = d->bn_shift;

  if (make_view_map(&p.map_a[0], d->in[0], 1 << p.lbw, 1 << p.lbh, 1 << p.lbn)) return 1;
  if (use_view1) {
    EB_REQUIRE(d->in[1].ptr, "eb200_conv2d: tap uses view 1 but in[1] is null");
    if (make_view_map(&p.map_a[1], d->in[1], 1 << p.lbw, 1 << p.lbh, 1 << p.lbn)) return 1;
  } else {
    p.map_a[1] = p.map_a[0];
  }
  // cluster size: share the weight tile across 2 / 4 consecutive pixel tiles when there is enough work
  int cluster = 1;
  {
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("EB200_CONV_CLUSTER");
      forced = e ? atoi(e) : 0;
    }
    const int want = forced > 0 ? forced : 1;   // measured: TMA multicast at cluster size <= 4 does not reduce L2->SM traffic
    for (int c = want; c > 1; c >>= 1) {
      if (!p.b_resident && block_n % (16 * c) == 0 && block_n >= 64 && tiles_m >= 4 * c) { cluster = c; break; }
    }
  }
  p.cluster = cluster;
  if (make_weight_map(&p.map_b, d->weight, d->cin_pad, d->cout_pad, d->weight_taps, block_n / cluster)) return 1;

  const int smem = conv_smem_bytes(block_n, stages, b_res_bytes, ksub);
  typedef void (*KernelFn)(const ConvParams);
  KernelFn fn = conv_tc_kernel<kGenericFlags>;
  switch (d->flags) {
    case 0: fn = conv_tc_kernel<0>; break;
    case kBias: fn = conv_tc_kernel<kBias>; break;
    case kBias | kRelu: fn = conv_tc_kernel<kBias | kRelu>; break;
    case kAuxAdd: fn = conv_tc_kernel<kAuxAdd>; break;
    case kStats: fn = conv_tc_kernel<kStats>; break;
    case kAuxMask | kStats | kStatsSum: fn = conv_tc_kernel<kAuxMask | kStats | kStatsSum>; break;
    default: break;
  }
  static KernelFn configured[16] = {};
  {
    int i = 0;
    for (; i < 16 && configured[i] && configured[i] != fn; ++i) {}
    if (i < 16 && !configured[i]) {
      EB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()));
      configured[i] = fn;
    }
  }
  const int groups = ceil_div(tiles_m, cluster) * p.tiles_c;
  int max_clusters = num_sms() / cluster;
  if (cluster == 4) max_clusters = 33;    // GPC granularity strands SMs at cluster size 4 (148 SMs: 132 usable)
  const int nclusters = groups < max_clusters ? groups : max_clusters;
  static long long* dbg_dev = nullptr;
  if (p.debug & 8) {
    if (!dbg_dev) { cudaMalloc(&dbg_dev, 8 * 64 * sizeof(long long)); }
    cudaMemset(dbg_dev, 0, 8 * 64 * sizeof(long long));
    p.dbg_buf = dbg_dev;
  }
  {
    void* kargs[1] = {&p};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(fn), dim3(nclusters * cluster), dim3(kThreads), smem,
                      static_cast<cudaStream_t>(stream), kargs, cluster, 4));
  }
  if (p.debug & 8) {
    static int printed = 0;
    long long h[8 * 64];
    cudaDeviceSynchronize();
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    if (printed++ % 23 == 3) {
      const long long t0 = h[0];
      printf("--- conv trace: Cout=%d block_n=%d taps=%d kblocks=%d stages=%d resident=%d\n", d->cout, block_n, d->taps, d->cin_pad / 64, stages, p.b_resident);
      for (int tl = 0; tl < 6; ++tl) {
        printf("tile %d: mma start %lld tempty_ok %lld |", tl, h[tl * 64] - t0, h[tl * 64 + 1] - t0);
        for (int i = 0; i < 12 && h[tl * 64 + 2 + 2 * i]; ++i) printf(" it%d full %lld commit %lld |", i, h[tl * 64 + 2 + 2 * i] - t0, h[tl * 64 + 3 + 2 * i] - t0);
        printf(" tfull-commit %lld || epi wait %lld got %lld chunks", h[tl * 64 + 30] - t0, h[tl * 64 + 32] - t0, h[tl * 64 + 33] - t0);
        for (int c = 0; c < 4 && h[tl * 64 + 34 + c]; ++c) printf(" %lld", h[tl * 64 + 34 + c] - t0);
        printf(" done %lld\n", h[tl * 64 + 40] - t0);
      }
      printf("tile2 it1 detail: after-wait %lld mma0 %lld mma1 %lld mma2 %lld mma3 %lld before-commit %lld after-commit %lld\n", h[7*64+9]-t0, h[7*64+0]-t0, h[7*64+1]-t0, h[7*64+2]-t0, h[7*64+3]-t0, h[7*64+8]-t0, h[7*64+10]-t0);
      printf("producer iteration starts:");
      for (int i = 0; i < 20; ++i) printf(" %lld", h[6 * 64 + i] - t0);
      printf("\n");
    }
  }
  return launch_check("conv_tc_kernel");
}

// dW[co][ci][3*tg + k] += ws[tg][co][ci][k]; ws = 0   (3x3 filters: the three tap groups were reduced into ws)
__global__ void __launch_bounds__(256) wgrad_ws_finish_kernel(float* __restrict__ dw, float* __restrict__ ws, int Cout,
                                                              int Cin) {
  const long long total = 9ll * Cin * Cout;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int t = static_cast<int>(i % 9);
    const long long r = i / 9;                       // co * Cin + ci
    const long long idx = ((static_cast<long long>(t / 3) * Cout * Cin) + r) * 3 + t % 3;
    dw[i] += ws[idx];
    ws[idx] = 0.f;
  }
}

extern "C" int eb200_conv2d_pair(const eb200_conv_desc* a, const eb200_conv_desc* b, void* stream) {
  EB_REQUIRE(a && b, "eb200_conv2d_pair: null argument");
  bool handled = false;
  if (a->taps == 3 && b->taps == 3 && a->cout >= 256 && a->out && b->out && a->weight && b->weight && a->in[0].ptr &&
      b->in[0].ptr && (a->out_sw % 8) == 0 && (b->out_sw % 8) == 0) {
    if (launch_conv3_dual(a, b, stream, &handled)) return 1;
  }
  if (handled) return 0;
  if (eb200_conv2d(a, stream)) return 1;
  return eb200_conv2d(b, stream);
}

extern "C" int eb200_conv2d_wgrad(const eb200_wgrad_desc* d, void* stream);
extern "C" int eb200_conv2d_wgrad_pair(const eb200_wgrad_desc* a, const eb200_wgrad_desc* b, void* stream) {
  EB_REQUIRE(a && b, "eb200_conv2d_wgrad_pair: null argument");
  bool handled = false;
  if (a->dw && b->dw && a->dy.ptr && b->dy.ptr && a->x[0].ptr && b->x[0].ptr && a->x[0].c >= 256) {
    if (launch_wgrad3_dual(a, b, stream, &handled)) return 1;
  }
  if (handled) return 0;
  if (eb200_conv2d_wgrad(a, stream)) return 1;
  return eb200_conv2d_wgrad(b, stream);
}

extern "C" int eb200_conv2d_wgrad(const eb200_wgrad_desc* d, void* stream) {
  EB_REQUIRE(d && d->dw && d->dy.ptr && d->x[0].ptr, "eb200_conv2d_wgrad: null argument");
  EB_REQUIRE(d->taps >= 1 && d->taps <= kMaxTaps, "eb200_conv2d_wgrad: taps=%d", d->taps);
  {
    bool handled = false;
    if (launch_wgrad3(d, stream, &handled)) return 1;
    if (handled) return 0;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->dy.n; p.H = d->dy.h; p.W = d->dy.w;
  p.Cout = d->dy.c; p.Cin = d->x[0].c;
  p.total_taps = d->taps;
  bool use_view1 = false;
  for (int t = 0; t < d->taps; ++t) {
    p.tap_view[t] = d->tap_view[t]; p.tap_dy[t] = d->tap_dy[t]; p.tap_dx[t] = d->tap_dx[t];
    use_view1 |= d->tap_view[t] == 1;
  }
  const int cin_r = ceil_div(p.Cin, 64) * 64;
  p.block_n = cin_r < (d->taps == 1 ? 256 : 128) ? cin_r : (d->taps == 1 ? 256 : 128);
  int tpi = 512 / p.block_n;
  if (tpi > 3) tpi = 3;
  if (tpi > d->taps) tpi = d->taps;
  p.taps = tpi;
  p.tap_groups = ceil_div(d->taps, tpi);
  choose_box(p.W, p.H, p.N, 6, &p.lbw, &p.lbh, &p.lbn);
  p.tiles_w = ceil_div(p.W, 1 << p.lbw);
  p.tiles_h = ceil_div(p.H, 1 << p.lbh);
  p.tiles_n = ceil_div(p.N, 1 << p.lbn);
  const int total_boxes = p.tiles_w * p.tiles_h * p.tiles_n;
  const int items = ceil_div(p.Cout, 128) * ceil_div(p.Cin, p.block_n) * p.tap_groups;
  int ksplit = (num_sms() + items / 2) / items;
  if (ksplit > ceil_div(total_boxes, 4)) ksplit = ceil_div(total_boxes, 4);
  if (ksplit < 1) ksplit = 1;
  // every CTA must own at least one box
  while (ksplit > 1 && ceil_div(total_boxes, ksplit) * (ksplit - 1) >= total_boxes) --ksplit;
  p.ksplit = ksplit;
  int stages = (smem_limit() - wgrad_smem_bytes(p.block_n, tpi, 0)) / wgrad_stage_bytes(p.block_n, tpi);
  if (stages > 6) stages = 6;
  EB_REQUIRE(stages >= 2, "eb200_conv2d_wgrad: not enough shared memory");
  p.stages = stages;
  p.dw = getenv("EB200_WGRAD_NOSTORE") ? nullptr : d->dw; p.dw_sco = d->dw_sco;
  p.vec_ok = (reinterpret_cast<uintptr_t>(d->dw) & 15) == 0; p.dw_sci = d->dw_sci; p.dw_st = d->dw_st;
  {
    const bool aligned = p.vec_ok && (p.Cin & 15) == 0 && (d->dw_sco & 3) == 0 && !getenv("EB200_WGRAD_NO_BULK");
    p.bulk = 0;
    if (aligned && d->taps == 1 && d->dw_sci == 1) p.bulk = 1;
    else if (aligned && d->taps == 3 && tpi == 3 && d->dw_st == 1 && d->dw_sci == 3) p.bulk = 2;
    else if (aligned && d->taps == 9 && tpi == 3 && d->dw_st == 1 && d->dw_sci == 9 && d->dw_sco == 9ll * p.Cin && d->ws &&
             (reinterpret_cast<uintptr_t>(d->ws) & 15) == 0 && d->ws_floats >= 9ll * p.Cin * p.Cout)
      p.bulk = 3;
    const long long need = 128ll * ((p.bulk == 1 ? 1 : 3) * p.block_n * 4 + 16);   // the epilogue re-uses the stage buffers
    if (p.bulk && static_cast<long long>(stages) * wgrad_stage_bytes(p.block_n, tpi) < need) p.bulk = 0;
    p.ws = d->ws;
  }

  if (make_view_map(&p.map_dy, d->dy, 1 << p.lbw, 1 << p.lbh, 1 << p.lbn)) return 1;
  if (make_view_map(&p.map_x[0], d->x[0], 1 << p.lbw, 1 << p.lbh, 1 << p.lbn)) return 1;
  if (use_view1) {
    EB_REQUIRE(d->x[1].ptr, "eb200_conv2d_wgrad: tap uses view 1 but x[1] is null");
    if (make_view_map(&p.map_x[1], d->x[1], 1 << p.lbw, 1 << p.lbh, 1 << p.lbn)) return 1;
  } else {
    p.map_x[1] = p.map_x[0];
  }
  const int smem = wgrad_smem_bytes(p.block_n, tpi, stages);
  static bool configured = false;
  if (!configured) {
    EB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()));
    configured = true;
  }
  {
    void* kargs[1] = {&p};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(wgrad_tc_kernel), dim3(items * ksplit), dim3(kWgThreads), smem,
                      static_cast<cudaStream_t>(stream), kargs, 1, 4));
  }
  if (launch_check("wgrad_tc_kernel")) return 1;
  if (p.bulk == 3) {
    const long long total = 9ll * p.Cin * p.Cout;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    wgrad_ws_finish_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d->dw, d->ws, p.Cout, p.Cin);
    return launch_check("wgrad_ws_finish_kernel");
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                   __nv_bfloat16* __restrict__ packed, int rows_pad, int cols_pad, int transpose,
                                   int co_off, int ci_off) {
  const int total = cout * cin * taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % taps;
    const int ci = (i / taps) % cin;
    const int co = i / (taps * cin);
    const int row = transpose ? ci + ci_off : co + co_off;
    const int col = transpose ? co + co_off : ci + ci_off;
    packed[(static_cast<size_t>(t) * rows_pad + row) * cols_pad + col] = __float2bfloat16(w[i]);
  }
}

extern "C" int eb200_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, void* packed, int cout_pad,
                                      int cin_pad, int transpose, int co_offset, int ci_offset, void* stream) {
  EB_REQUIRE(w && packed, "eb200_pack_conv_weight: null argument");
  const int taps = kh * kw;
  const int rows = transpose ? cin + ci_offset : cout + co_offset;
  const int cols = transpose ? cout + co_offset : cin + ci_offset;
  EB_REQUIRE(rows <= cout_pad && cols <= cin_pad, "eb200_pack_conv_weight: block %dx%d exceeds padded %dx%d", rows,
             cols, cout_pad, cin_pad);
  const int total = cout * cin * taps;
  int blocks = ceil_div(total, 256);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  pack_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, taps, static_cast<__nv_bfloat16*>(packed), cout_pad, cin_pad, transpose, co_offset, ci_offset);
  return launch_check("pack_weight_kernel");
}

// ------------------------------------------------------------------------------------------------
// All weights of the model in ONE launch (a training step changes every weight).  Block b handles the 32 x 32
// (co, ci) tile (co0, ci0) = (block_start[b] & 0xffff, block_start[b] >> 16) * 32 of entry block_entry[b], all taps:
// the fp32 source rows are read coalesced into shared memory, then both bf16 layouts are written as 64-byte row
// segments (ci contiguous for the forward layout, co contiguous for the transposed one).
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const eb200_pack_entry* __restrict__ entries,
                                                                    const int* __restrict__ block_entry,
                                                                    const int* __restrict__ block_start) {
  __shared__ float tile[32][32 * 9 + 1];
  const eb200_pack_entry e = entries[block_entry[blockIdx.x]];
  const int co0 = (block_start[blockIdx.x] & 0xffff) * 32, ci0 = (block_start[blockIdx.x] >> 16) * 32;
  const int nco = min(32, e.cout - co0), nci = min(32, e.cin - ci0);
  const int rowlen = nci * e.taps;                       // contiguous floats of one co row of the tile
  for (int i = threadIdx.x; i < nco * rowlen; i += 256) {
    const int r = i / rowlen, c = i - r * rowlen;
    tile[r][c] = __ldg(e.w + (static_cast<size_t>(co0 + r) * e.cin + ci0) * e.taps + c);
  }
  __syncthreads();
  __nv_bfloat16* fwd = static_cast<__nv_bfloat16*>(e.fwd);
  __nv_bfloat16* bwd = static_cast<__nv_bfloat16*>(e.bwd);
  const int total = e.taps * 32 * 32;
  for (int i = threadIdx.x; i < total; i += 256) {       // forward layout: ci fastest
    const int ci = i & 31, co = (i >> 5) & 31, t = i >> 10;
    if (co < nco && ci < nci)
      fwd[(static_cast<size_t>(t) * e.fwd_rows + co0 + co + e.co_off) * e.fwd_cols + ci0 + ci + e.ci_off] =
          __float2bfloat16(tile[co][ci * e.taps + t]);
  }
  if (bwd) {
    for (int i = threadIdx.x; i < total; i += 256) {     // transposed layout: co fastest
      const int co = i & 31, ci = (i >> 5) & 31, t = i >> 10;
      if (co < nco && ci < nci)
        bwd[(static_cast<size_t>(t) * e.bwd_rows + ci0 + ci + e.ci_off) * e.bwd_cols + co0 + co + e.co_off] =
            __float2bfloat16(tile[co][ci * e.taps + t]);
    }
  }
}

extern "C" int eb200_pack_conv_weights_batched(const eb200_pack_entry* entries_dev, const int* block_entry_dev,
                                               const int* block_start_dev, int nblocks, void* stream) {
  EB_REQUIRE(entries_dev && block_entry_dev && block_start_dev && nblocks > 0,
             "eb200_pack_conv_weights_batched: bad argument");
  pack_weights_batched_kernel<<<nblocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(entries_dev, block_entry_dev,
                                                                                      block_start_dev);
  return launch_check("pack_weights_batched_kernel");
}
