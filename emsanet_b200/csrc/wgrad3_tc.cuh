// Weight gradient of the 3-tap 1-D stride-1 convolutions (3x1 / 1x3 of NonBottleneck1D, MT/model/block.py:174-190;
// aten convolution_backward, weight part, entered at main.py:598) on tcgen05 — "halo" formulation.
//
//   dW[co][ci][t] += sum_pixels dY[p, co] * X[p + off_t, ci]          GEMM: M = co (128), N = ci (BN), K = pixels
//
// Both operands are [pixel][channel] in HBM, i.e. MN-major for this GEMM; a pipeline stage covers a box of
// F x S = 64 pixels of one image (F along the non-tap axis, S along the tap axis).  The generic kernel loads the X box
// once per tap; here ONE box with a 2-pixel halo along the tap axis — (64 ch, F, S+2) — is loaded per 64-channel group
// and the three taps are three UMMA B-descriptors into it, (off_t + 1) * F rows apart (a multiple of the 8-row swizzle
// atom since F >= 8).  Out-of-image rows are the TMA zero fill (== zero padding; zero dY rows add nothing).
// The three taps accumulate side by side in TMEM (3 x BN columns) and leave as 16-byte vector reductions straight into
// the parameter's .grad in the reference layout [Cout][Cin][3].  Control loops are warp-uniform (one lane issues).
#pragma once
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace eb {

struct Wgrad3Params {
  CUtensorMap map_dy;        // dims (C, fast, slow, N), box (64, F, S, 1)
  CUtensorMap map_x;         // dims (C, fast, slow, N), box (64, F, S + 2, 1)
  int N, ext_f, ext_s;
  int Cin, Cout;
  int lgF;                   // F = 1 << lgF in {8, 16, 32}; S = 64 >> lgF
  int tiles_f, tiles_s;
  int total_boxes;           // tiles_f * tiles_s * N
  int ci_tiles, ksplit;
  int tap_row[3];            // (offset of tap t + 1) * F
  int stages;
  int x_sub_bytes;           // (S + 2) * F * 128
  float* dw;                 // fp32 [Cout][Cin][3]
  long long dw_sco;          // = Cin * 3
};

constexpr int kWg3Threads = 192;   // 2 control warps + 4 epilogue warps
constexpr int kWg3DySub = 64 * 128;

// DUAL: even CTAs work on pa, odd CTAs on pb (two independent problems of identical geometry in one launch)
// CL = 2: the two CTAs of a cluster hold adjacent K-splits of the SAME output tile; before the bulk reduce-add they sum
// their partial tiles through distributed shared memory (each CTA finishes 64 of the 128 rows), which halves the
// split-K reduction traffic the L2 has to absorb (148 partial 128 x 384 fp32 tiles per launch otherwise).
// (the single-problem kernel takes ONE parameter block, see conv3_tc.cuh)
template <int BN, int CL>
__device__ __forceinline__ void wgrad3_tc_body(const Wgrad3Params& p, int bid) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int NSUB = BN / 64;
  const uint32_t stage_bytes = 2 * kWg3DySub + NSUB * p.x_sub_bytes;
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (8 + s); };
  const uint32_t tfull_bar = bar_base + 8u * 16;
  const uint32_t tmem_slot = bar_base + 8u * 17;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // work item: blockIdx.x = (co_tile * ci_tiles + ci_tile) * ksplit + ks
  const int ks = bid % p.ksplit; bid /= p.ksplit;
  const int ci_tile = bid % p.ci_tiles;
  const int co_tile = bid / p.ci_tiles;
  const int per = (p.total_boxes + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(p.total_boxes, kb0 + per);
  const int nk = max(0, kb1 - kb0);
  const bool two_dy = co_tile * 128 + 64 < p.Cout;   // second 64-channel group of dY exists

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_dy);
    tma_prefetch_desc(&p.map_x);
    for (int s = 0; s < 8; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();   // the set-up above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int S = 64 >> p.lgF;
  const int tiles_fs = p.tiles_f * p.tiles_s;

  if (nk > 0) {
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      const uint32_t tx_bytes = (two_dy ? 2 : 1) * kWg3DySub + NSUB * p.x_sub_bytes;
      uint32_t s = 0, ph = 0;
      int n = kb0 / tiles_fs, rem = kb0 - n * tiles_fs;
      int ts = rem / p.tiles_f, tf = rem - ts * p.tiles_f;
      for (int i = 0; i < nk; ++i) {
        const int f0 = tf << p.lgF, s0 = ts * S;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (lane == 0) {
          const uint32_t sa = smem_base + s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), tx_bytes);
          tma_load_4d(sa, &p.map_dy, full_bar(s), co_tile * 128, f0, s0, n);
          if (two_dy) tma_load_4d(sa + kWg3DySub, &p.map_dy, full_bar(s), co_tile * 128 + 64, f0, s0, n);
#pragma unroll
          for (int j = 0; j < NSUB; ++j)
            tma_load_4d(sa + 2 * kWg3DySub + j * p.x_sub_bytes, &p.map_x, full_bar(s), ci_tile * BN + j * 64, f0,
                        s0 - 1, n);
        }
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
        if (++tf == p.tiles_f) { tf = 0; if (++ts == p.tiles_s) { ts = 0; ++n; } }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < nk; ++i) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + s * stage_bytes;
          const uint32_t sx = sa + 2 * kWg3DySub;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 16 pixels (2 swizzle atoms of 8 rows) per UMMA
              const uint64_t adesc = make_smem_desc(sa + k * 2048, kWg3DySub, 1024);
              const uint64_t bdesc = make_smem_desc(sx + p.tap_row[t] * 128 + k * 2048, p.x_sub_bytes, 1024);
              umma_bf16(tmem_base + t * BN, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(s));
        }
        if (++s == static_cast<uint32_t>(p.stages)) { s = 0; ph ^= 1u; }
      }
      if (lane == 0) umma_commit(tfull_bar);
      __syncwarp();
    } else {
      // ---------------------------------------------------------------- epilogue: TMEM -> smem row -> bulk reduce-add
      // Thread r owns accumulator row r (= output channel co): 3 x BN fp32 that are ONE contiguous run
      // dW[co][ci_tile*BN .. +BN][0..2] of the reference layout.  It interleaves the three taps into its own smem row
      // (the pipeline buffers are dead once tfull fires) and hands the run to the TMA as a single reduce-add:
      // the L2 adds whole lines instead of 96 scattered 16-byte atomics per thread.
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const int co = co_tile * 128 + row;
      constexpr uint32_t kRowBytes = 3 * BN * 4;
      constexpr uint32_t kRowPitch = kRowBytes + 16;          // +16 B: rows start in different banks
      const uint32_t srow = smem_base + row * kRowPitch;
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < BN / 16; ++g) {
        uint32_t v[3][16];
        tmem_ld16(t_row + 0 * BN + g * 16, v[0]);
        tmem_ld16(t_row + 1 * BN + g * 16, v[1]);
        tmem_ld16(t_row + 2 * BN + g * 16, v[2]);
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
      }
      if (CL == 1) {
        fence_proxy_async();   // my generic-proxy smem writes -> visible to the bulk (async proxy) read
        if (co < p.Cout && p.dw != nullptr) {
          bulk_reduce_add_f32(p.dw + co * p.dw_sco + static_cast<long long>(ci_tile) * BN * 3, srow, kRowBytes);
          bulk_commit_group();
          bulk_wait_group_read0();   // the smem row must outlive the read
        }
      }
    }
  } else if (CL == 2 && warp >= 2) {   // (no pixel boxes: contribute a zero tile to the pair)
    constexpr uint32_t kPitch0 = 3 * BN * 4 + 16;
    const uint32_t srow0 = smem_base + ((warp & 3) * 32 + lane) * kPitch0;
    for (uint32_t o = 0; o < 3 * BN * 4; o += 16)
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(srow0 + o), "r"(0u) : "memory");
  }
  if (CL == 2) {
    // pair reduction through DSMEM: rank q finishes rows 64q .. 64q+63 (its own partial + the peer's), two threads per row
    constexpr uint32_t kRowBytes2 = 3 * BN * 4;
    constexpr uint32_t kPitch2 = kRowBytes2 + 16;
    __syncthreads();
    cluster_sync_all();                        // both partial tiles are complete in shared memory
    const uint32_t rank = cluster_ctarank();
    if (warp >= 2) {
      const int e = (warp - 2) * 32 + lane;    // 0..127
      const int row = static_cast<int>(rank) * 64 + (e & 63);
      const uint32_t half_off = static_cast<uint32_t>(e >> 6) * (kRowBytes2 / 2);
      const uint32_t lrow = smem_base + row * kPitch2 + half_off;
      const uint32_t rrow = mapa_shared(lrow, rank ^ 1u);
#pragma unroll 4
      for (uint32_t o = 0; o < kRowBytes2 / 2; o += 16) {
        float4 a;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(lrow + o));
        const float4 b = ld_dsmem_f4(rrow + o);
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lrow + o), "f"(a.x + b.x), "f"(a.y + b.y),
                     "f"(a.z + b.z), "f"(a.w + b.w) : "memory");
      }
      fence_proxy_async();
      named_bar_sync(1, 128);                  // both halves of every row are summed
      if (e < 64) {
        const int co = co_tile * 128 + row;
        if (co < p.Cout && p.dw != nullptr) {
          bulk_reduce_add_f32(p.dw + co * p.dw_sco + static_cast<long long>(ci_tile) * BN * 3,
                              smem_base + row * kPitch2, kRowBytes2);
          bulk_commit_group();
          bulk_wait_group_read0();
        }
      }
    }
    cluster_sync_all();                        // the peer has finished reading my partial tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, int CL = 1>
__global__ void __launch_bounds__(kWg3Threads, 1) wgrad3_tc_kernel(const __grid_constant__ Wgrad3Params p) {
  wgrad3_tc_body<BN, CL>(p, static_cast<int>(blockIdx.x));
}

template <int BN>
__global__ void __launch_bounds__(kWg3Threads, 1) wgrad3_tc_dual_kernel(const __grid_constant__ Wgrad3Params pa,
                                                                        const __grid_constant__ Wgrad3Params pb) {
  if (blockIdx.x & 1)
    wgrad3_tc_body<BN, 1>(pb, static_cast<int>(blockIdx.x >> 1));
  else
    wgrad3_tc_body<BN, 1>(pa, static_cast<int>(blockIdx.x >> 1));
}

}  // namespace eb
