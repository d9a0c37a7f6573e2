// HBM-bound kernels of the EMSANet path (sm_100a): batch-norm statistics/apply/backward, SE fusion,
// max/adaptive pooling, bilinear and learned (nearest + depthwise 3x3) upsampling, layout conversion,
// head activations, scene linear layer.  All activations NHWC bf16 with C % 8 == 0 so that every thread
// moves 16 bytes per access (8 channels); per-channel parameters and reductions are fp32.
// Grids are sized as multiples of the SM count with grid-stride loops.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <cstring>

#include "../../include/emsanet_b200.h"
#include "common.h"

namespace eb {

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void cvt8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

static inline int grid_for(long long work_items, int threads, int waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------------------------------
// per-(n, c) / per-c reductions: a block walks a contiguous pixel range of ONE image; threads are laid
// out (pixel_lane, c8) with the channel group fastest, so global loads are coalesced and each thread
// keeps its 8 channels in registers; the pixel_lanes are combined through shared memory at the end.
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NV][8], int c8, int C8, int plane, int planes,
                                                     float* smem /* [planes][NV][C8*8] */) {
  // all threads of the block call this; threads with plane >= planes are idle padding.
  // result: smem[v * C + c] (plane 0), valid for the threads with plane == 0 afterwards.
  const int C = C8 * 8;
  if (plane < planes) {
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) smem[(plane * NV + v) * C + c8 * 8 + j] = acc[v][j];
  }
  __syncthreads();
  if (plane == 0) {
    for (int pl = 1; pl < planes; ++pl)
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) smem[v * C + c8 * 8 + j] += smem[(pl * NV + v) * C + c8 * 8 + j];
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm (reference: nn.BatchNorm2d via MT/model/normalization.py:30-31; eps 1e-5, momentum 0.1)
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(float* __restrict__ stats, float count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s = stats[c], q = stats[C + c];
  stats[c] = 0.f;       // buffer is reusable for the next step
  stats[C + c] = 0.f;
  const double mean = s / count;
  double var = q / count - mean * mean;   // biased variance, used for normalisation
  if (var < 0) var = 0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  mean_out[c] = static_cast<float>(mean);
  rstd_out[c] = rstd;
  if (running_mean) {   // track_running_stats: unbiased variance in the running estimate
    const double unb = count > 1.f ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unb);
  }
}

// y = relu?( (x*scale+shift) * drop[n,c] + res_pre ) + res_post ; optional per-(n,c) sum of y (SE squeeze)
struct BnApplyArgs {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* scale;
  const float* shift;
  const float* drop;             // [N][C] or null
  const __nv_bfloat16* res_pre;  // or null
  const __nv_bfloat16* res_post; // or null
  float* gap;                    // [N][C] or null (atomic accumulate)
  int N, HW, C;
  int y_cs, y_coff;              // channel pitch / offset of y (concat fusion)
  int relu;
  int chunks;                    // blocks per image
  // train mode with the finalize folded in (stats != null): every block derives the affine of its channels from the raw
  // sums; block 0 also publishes scale/shift/mean/rstd for the backward pass and updates the running statistics
  const float* stats;            // [2C] sum, sum of squares (left untouched: the caller zeroes its arena once per step)
  const float* gamma;
  const float* beta;
  float* running_mean;           // or null
  float* running_var;
  float* scale_out;
  float* shift_out;
  float* mean_out;
  float* rstd_out;
  float count, eps, momentum;
};

__global__ void __launch_bounds__(256, 2) bn_apply_kernel(BnApplyArgs a) {
  extern __shared__ float red[];
  pdl_launch_dependents();
  pdl_wait();
  const int C8 = a.C >> 3;
  const int planes = 256 / C8;              // pixel lanes per block
  const int c8 = threadIdx.x % C8;
  const int plane = threadIdx.x / C8;
  const int n = blockIdx.x / a.chunks;
  const int chunk = blockIdx.x - n * a.chunks;
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per;
  const int p1 = min(a.HW, p0 + per);
  float sc[8], sh[8], dr[8];
  if (a.stats != nullptr) {
    float su[8], sq[8], ga[8], be[8];
    load8f(a.stats + c8 * 8, su);
    load8f(a.stats + a.C + c8 * 8, sq);
    load8f(a.gamma + c8 * 8, ga);
    load8f(a.beta + c8 * 8, be);
    const bool publish = blockIdx.x == 0 && plane == 0;
    const float inv_cnt = 1.f / a.count;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // fp32 with the square folded into one fma: var = E[x^2] - mean^2 loses ~6e-8 * (1 + mean^2/var) relative,
      // far below the bf16 activations it normalises (fp64 here would put ~40 slow DFMA/DDIV chains in front of
      // every block of a bandwidth-bound kernel)
      const float mean = su[j] * inv_cnt;
      const float var = fmaxf(fmaf(-mean, mean, sq[j] * inv_cnt), 0.f);   // biased variance, used for normalisation
      const float rstd = 1.f / sqrtf(var + a.eps);
      sc[j] = ga[j] * rstd;
      sh[j] = fmaf(-mean, sc[j], be[j]);
      if (publish) {
        const int c = c8 * 8 + j;
        a.scale_out[c] = sc[j];
        a.shift_out[c] = sh[j];
        a.mean_out[c] = mean;
        a.rstd_out[c] = rstd;
        if (a.running_mean) {   // track_running_stats: unbiased variance in the running estimate
          const float unb = a.count > 1.f ? var * (a.count / (a.count - 1.f)) : var;
          a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * mean;
          a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * unb;
        }
      }
    }
  } else {
    load8f(a.scale + c8 * 8, sc);
    load8f(a.shift + c8 * 8, sh);
  }
  if (a.drop) load8f(a.drop + static_cast<size_t>(n) * a.C + c8 * 8, dr);
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  constexpr int U = 4;                      // pixels in flight per thread (memory-level parallelism)
  if (plane < planes) {
    for (int pb = p0 + plane; pb < p1; pb += planes * U) {
      uint4 rv[U], rr0[U], rr1[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * planes;
        ok[u] = p < p1;
        if (ok[u]) {
          const size_t pix = static_cast<size_t>(n) * a.HW + p;
          rv[u] = ldg16(a.x + pix * a.C + c8 * 8);
          if (a.res_pre) rr0[u] = ldg16(a.res_pre + pix * a.C + c8 * 8);
          if (a.res_post) rr1[u] = ldg16(a.res_post + pix * a.C + c8 * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const size_t pix = static_cast<size_t>(n) * a.HW + pb + u * planes;
        float v[8], r0[8], r1[8];
        cvt8(rv[u], v);
        if (a.res_pre) cvt8(rr0[u], r0);
        if (a.res_post) cvt8(rr1[u], r1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = fmaf(v[j], sc[j], sh[j]);
          if (a.drop) t *= dr[j];
          if (a.res_pre) t += r0[j];
          if (a.relu) t = fmaxf(t, 0.f);
          if (a.res_post) t += r1[j];
          v[j] = t;
        }
        store8(a.y + pix * a.y_cs + a.y_coff + c8 * 8, v);
        if (a.gap) {
          // squeeze statistics are taken from the stored (bf16-rounded) activations
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[0][j] += __bfloat162float(__float2bfloat16(v[j]));
        }
      }
    }
  }
  if (a.gap) {
    block_channel_reduce<1>(acc, c8, C8, plane, planes, red);
    if (plane == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(a.gap + static_cast<size_t>(n) * a.C + c8 * 8 + j, red[c8 * 8 + j]);
    }
  }
}

// g = dy * relu_mask * drop ; sums[0:C] += sum g ; sums[C:2C] += sum g * xhat
struct BnBwdArgs {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* x;        // raw conv output (BN input)
  const __nv_bfloat16* mask_src; // tensor whose sign gives the ReLU mask, or null
  const float* drop;             // [N][C] or null
  const float* mean;
  const float* rstd;
  const float* scale;            // gamma*rstd (for recomputing the mask)
  const float* shift;
  const float* gamma;
  float* sums;                   // [2C]
  __nv_bfloat16* dx;             // apply only
  __nv_bfloat16* dres;           // apply only: dy*relu_mask (gradient of the pre-ReLU sum), or null
  int N, HW, C;
  int dy_cs, dy_coff;            // channel pitch/offset of dy (slices of wider gradient tensors)
  int relu_mode;                 // 0 none, 1 mask_src > 0, 2 recompute x*scale+shift > 0
  int chunks;
  float inv_count;
  // replica mode (replicas > 0, reduce kernel): block sums go to sums[blockIdx % replicas][2C] with atomics (few blocks
  // per address); the last block to finish folds the replicas into sums[replicas][2C] = (sum g, sum g*xhat), adds
  // dgamma / dbeta; counter = (unsigned*)(sums + (replicas + 1) * 2C).  Everything ZERO on entry.
  int replicas;
  float* dgamma;
  float* dbeta;
  int raw_sums;                  // apply kernel: sums = (sum g, sum g*x) as left by a fused conv epilogue (EB200_BN_BWD)
  int fused;                     // reduce kernel (replica mode): continue with the apply phase after a grid-wide barrier —
                                 // every block re-reads its own pixel range (L1/L2 hot) and writes dx / dres.  Needs
                                 // all blocks co-resident (the host checks the occupancy).
};

// finishes g from already loaded operands: g = dy * relu_mask (-> gres) * drop
__device__ __forceinline__ void bn_bwd_g(const BnBwdArgs& a, const float (&sc)[8], const float (&sh)[8],
                                         const float (&dr)[8], const float (&xv)[8], const float (&m)[8],
                                         float (&g)[8], float (&gres)[8]) {
  if (a.relu_mode == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = m[j] > 0.f ? g[j] : 0.f;
  } else if (a.relu_mode == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = fmaf(xv[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) gres[j] = g[j];
  if (a.drop) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= dr[j];
  }
}

__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(BnBwdArgs a) {
  extern __shared__ float red[];
  pdl_launch_dependents();
  pdl_wait();
  const int C8 = a.C >> 3;
  const int planes = 256 / C8;
  const int c8 = threadIdx.x % C8;
  const int plane = threadIdx.x / C8;
  const int n = blockIdx.x / a.chunks;
  const int chunk = blockIdx.x - n * a.chunks;
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per, p1 = min(a.HW, p0 + per);
  float sc[8], sh[8], dr[8];
  if (a.relu_mode == 2) {
    load8f(a.scale + c8 * 8, sc);
    load8f(a.shift + c8 * 8, sh);
  }
  if (a.drop) load8f(a.drop + static_cast<size_t>(n) * a.C + c8 * 8, dr);
  float acc[2][8];   // sum g and sum g*x (xhat is applied in the combine kernel: sum g*xhat = rstd*(sum gx - mean*sum g))
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  constexpr int U = 4;
  if (plane < planes) {
    for (int pb = p0 + plane; pb < p1; pb += planes * U) {
      uint4 rx[U], rg[U], rm[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * planes;
        ok[u] = p < p1;
        if (ok[u]) {
          const size_t pix = static_cast<size_t>(n) * a.HW + p;
          rx[u] = ldg16(a.x + pix * a.C + c8 * 8);
          rg[u] = ldg16(a.dy + pix * a.dy_cs + a.dy_coff + c8 * 8);
          if (a.relu_mode == 1) rm[u] = ldg16(a.mask_src + pix * a.C + c8 * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        float xv[8], g[8], m[8], gres[8];
        cvt8(rx[u], xv);
        cvt8(rg[u], g);
        if (a.relu_mode == 1) cvt8(rm[u], m);
        bn_bwd_g(a, sc, sh, dr, xv, m, g, gres);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += g[j];
          acc[1][j] = fmaf(g[j], xv[j], acc[1][j]);
        }
      }
    }
  }
  block_channel_reduce<2>(acc, c8, C8, plane, planes, red);
  if (a.replicas > 0) {
    if (plane == 0) {
      float* dst = a.sums + static_cast<size_t>(blockIdx.x % a.replicas) * 2 * a.C;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dst + c8 * 8 + j, red[c8 * 8 + j]);
        atomicAdd(dst + a.C + c8 * 8 + j, red[a.C + c8 * 8 + j]);
      }
    }
    // last block done: fold the replicas (threadfence + counter: the standard "last block" reduction)
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned* counter = reinterpret_cast<unsigned*>(a.sums + static_cast<size_t>(a.replicas + 1) * 2 * a.C);
      is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      float* folded = a.sums + static_cast<size_t>(a.replicas) * 2 * a.C;
      for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
        float sg = 0.f, sgx = 0.f;
        for (int r = 0; r < a.replicas; ++r) {
          sg += __ldcg(a.sums + static_cast<size_t>(r) * 2 * a.C + c);
          sgx += __ldcg(a.sums + static_cast<size_t>(r) * 2 * a.C + a.C + c);
        }
        const float v = a.rstd[c] * (sgx - a.mean[c] * sg);   // sum g*xhat
        folded[c] = sg;
        folded[a.C + c] = v;
        a.dbeta[c] += sg;
        a.dgamma[c] += v;
      }
      if (a.fused) {   // release the grid: the folded sums are complete
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
          unsigned* flag = reinterpret_cast<unsigned*>(a.sums + static_cast<size_t>(a.replicas + 1) * 2 * a.C) + 1;
          atomicExch(flag, 1u);
        }
      }
    }
    if (a.fused) {
      // grid-wide barrier (all blocks are resident: checked by the host), then the apply phase on this block's pixels
      if (threadIdx.x == 0) {
        volatile unsigned* flag =
            reinterpret_cast<volatile unsigned*>(a.sums + static_cast<size_t>(a.replicas + 1) * 2 * a.C) + 1;
        unsigned spins = 0;
        while (*flag == 0u) {
          __nanosleep(64);
          if (++spins > (1u << 24)) asm volatile("trap;");   // never hang the GPU on a protocol error
        }
      }
      __syncthreads();
      __threadfence();
      if (plane < planes) {
        const float* folded = a.sums + static_cast<size_t>(a.replicas) * 2 * a.C;
        float k0[8], k1[8], k2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // dx = gamma*rstd*(g - s0/n - xhat*s1/n) = k0*g - k1 - x*k2
          const int c = c8 * 8 + j;
          const float s0 = __ldcg(folded + c), s1 = __ldcg(folded + a.C + c);
          const float rs = a.rstd[c], mu = a.mean[c];
          k0[j] = a.gamma[c] * rs;
          const float t2 = k0[j] * s1 * a.inv_count * rs;
          k1[j] = k0[j] * s0 * a.inv_count - mu * t2;
          k2[j] = t2;
        }
        for (int pb = p0 + plane; pb < p1; pb += planes * U) {
          uint4 rx[U], rg[U], rm[U];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int p = pb + u * planes;
            ok[u] = p < p1;
            if (ok[u]) {
              const size_t pix = static_cast<size_t>(n) * a.HW + p;
              rx[u] = ldg16(a.x + pix * a.C + c8 * 8);
              rg[u] = ldg16(a.dy + pix * a.dy_cs + a.dy_coff + c8 * 8);
              if (a.relu_mode == 1) rm[u] = ldg16(a.mask_src + pix * a.C + c8 * 8);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const size_t pix = static_cast<size_t>(n) * a.HW + pb + u * planes;
            float xv[8], g[8], m[8], gres[8], o[8];
            cvt8(rx[u], xv);
            cvt8(rg[u], g);
            if (a.relu_mode == 1) cvt8(rm[u], m);
            bn_bwd_g(a, sc, sh, dr, xv, m, g, gres);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = k0[j] * g[j] - k1[j] - xv[j] * k2[j];
            store8(a.dx + pix * a.C + c8 * 8, o);
            if (a.dres) store8(a.dres + pix * a.C + c8 * 8, gres);
          }
        }
      }
    }
  } else if (plane == 0) {
    // per-block partials (no atomics: hundreds of blocks adding into the same 2C floats serialise in L2)
    float* dst = a.sums + static_cast<size_t>(blockIdx.x) * 2 * a.C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dst[c8 * 8 + j] = red[c8 * 8 + j];
      dst[a.C + c8 * 8 + j] = red[a.C + c8 * 8 + j];
    }
  }
}

// Per channel c: sg = sum_b partials[b][c], sgx = sum_b partials[b][C+c];
// sums[c] = sg, sums[C+c] = rstd*(sgx - mean*sg) (= sum g*xhat); dbeta += sums[c]; dgamma += sums[C+c].
// Block = 32 channels x 8 block-groups.
__global__ void __launch_bounds__(256) bn_bwd_combine_kernel(const float* __restrict__ partials, int nblocks,
                                                             float* __restrict__ sums, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, int C) {
  __shared__ float red[2][8][33];
  const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + col;
  float sg = 0.f, sgx = 0.f;
  if (c < C) {
    for (int b = grp; b < nblocks; b += 8) {
      sg += partials[static_cast<size_t>(b) * 2 * C + c];
      sgx += partials[static_cast<size_t>(b) * 2 * C + C + c];
    }
  }
  red[0][grp][col] = sg;
  red[1][grp][col] = sgx;
  __syncthreads();
  if (grp == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) { t0 += red[0][g][col]; t1 += red[1][g][col]; }
    const float v = rstd[c] * (t1 - mean[c] * t0);
    sums[c] = t0;
    sums[C + c] = v;
    dbeta[c] += t0;
    dgamma[c] += v;
  }
}

// same (n, chunk) x (plane, c8) decomposition as the reduce kernel so the per-channel vectors are loaded once
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(BnBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int C8 = a.C >> 3;
  const int planes = 256 / C8;
  const int c8 = threadIdx.x % C8;
  const int plane = threadIdx.x / C8;
  const int n = blockIdx.x / a.chunks;
  const int chunk = blockIdx.x - n * a.chunks;
  const int per = (a.HW + a.chunks - 1) / a.chunks;
  const int p0 = chunk * per, p1 = min(a.HW, p0 + per);
  if (plane >= planes) return;
  float sc[8], sh[8], dr[8], k0[8], k1[8], k2[8];
  if (a.relu_mode == 2) {
    load8f(a.scale + c8 * 8, sc);
    load8f(a.shift + c8 * 8, sh);
  }
  {
    float ga[8], s0[8], s1[8], mu[8], rs[8];
    load8f(a.gamma + c8 * 8, ga);
    load8f(a.mean + c8 * 8, mu);
    load8f(a.rstd + c8 * 8, rs);
    load8f(a.sums + c8 * 8, s0);
    load8f(a.sums + a.C + c8 * 8, s1);
    if (a.raw_sums) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s1[j] = rs[j] * (s1[j] - mu[j] * s0[j]);   // sum g*xhat
      if (blockIdx.x == 0 && plane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a.dbeta[c8 * 8 + j] += s0[j];
          a.dgamma[c8 * 8 + j] += s1[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // dx = gamma*rstd*(g - s0/n - xhat*s1/n) = k0*g - k1 - x*k2
      k0[j] = ga[j] * rs[j];
      const float t2 = k0[j] * s1[j] * a.inv_count * rs[j];
      k1[j] = k0[j] * s0[j] * a.inv_count - mu[j] * t2;
      k2[j] = t2;
    }
  }
  if (a.drop) load8f(a.drop + static_cast<size_t>(n) * a.C + c8 * 8, dr);
  constexpr int U = 4;
  for (int pb = p0 + plane; pb < p1; pb += planes * U) {
    uint4 rx[U], rg[U], rm[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * planes;
      ok[u] = p < p1;
      if (ok[u]) {
        const size_t pix = static_cast<size_t>(n) * a.HW + p;
        rx[u] = ldg16(a.x + pix * a.C + c8 * 8);
        rg[u] = ldg16(a.dy + pix * a.dy_cs + a.dy_coff + c8 * 8);
        if (a.relu_mode == 1) rm[u] = ldg16(a.mask_src + pix * a.C + c8 * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const size_t pix = static_cast<size_t>(n) * a.HW + pb + u * planes;
      float xv[8], g[8], m[8], gres[8], o[8];
      cvt8(rx[u], xv);
      cvt8(rg[u], g);
      if (a.relu_mode == 1) cvt8(rm[u], m);
      bn_bwd_g(a, sc, sh, dr, xv, m, g, gres);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = k0[j] * g[j] - k1[j] - xv[j] * k2[j];
      store8(a.dx + pix * a.C + c8 * 8, o);
      if (a.dres) store8(a.dres + pix * a.C + c8 * 8, gres);
    }
  }
}

// dgamma += sums[C:2C], dbeta += sums[0:C]; sums zeroed for reuse
__global__ void bn_bwd_param_kernel(float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] += sums[c];
  dgamma[c] += sums[C + c];
  sums[c] = 0.f;
  sums[C + c] = 0.f;
}

// out[c] += sum over pixels of x[p][cs*p + coff + c]   (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                     long long P, int C, int cs, int coff) {
  extern __shared__ float red[];
  const int C8 = C >> 3;
  const int planes = 256 / C8;
  const int c8 = threadIdx.x % C8, plane = threadIdx.x / C8;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  if (plane < planes) {
    for (long long p = static_cast<long long>(blockIdx.x) * planes + plane; p < P;
         p += static_cast<long long>(gridDim.x) * planes) {
      float v[8];
      load8(x + p * cs + coff + c8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] += v[j];
    }
  }
  block_channel_reduce<1>(acc, c8, C8, plane, planes, red);
  if (plane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(out + c8 * 8 + j, red[c8 * 8 + j]);
  }
}

// ------------------------------------------------------------------------------------------------
// Stem: im2col of the 7x7 stride-2 pad-3 window (MT/model/backbone/resnet.py:64-65) so the stem runs
// as a 1x1 implicit GEMM on the tensor cores.  in: fp32 NCHW; out: bf16 [N,Ho,Wo,Kpad], k = c*49+ky*7+kx
// (the reference weight's own flattening), zero for k >= Cin*49.
// ------------------------------------------------------------------------------------------------
// One block per output row (n, ho): the 7 input rows x Cin channels it needs are staged once in shared memory as bf16
// (coalesced fp32 reads, zero outside the image, 3-pixel zero halo left and right); a thread owns a fixed group of 8
// consecutive k = (c, ky, kx) — its 8 smem offsets are computed once — and walks the output pixels of the row:
// 8 two-byte smem reads and one 16-byte store per item instead of ~250 instructions of index arithmetic.
__global__ void __launch_bounds__(256) im2col_stem_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                          int N, int Cin, int H, int W, int Ho, int Wo, int Kpad) {
  extern __shared__ __nv_bfloat16 rows_s[];   // [Cin][7][W + 6]
  const int K8 = Kpad >> 3;
  const int K = Cin * 49;
  const int Wp = W + 6;
  const int ho = blockIdx.x, n = blockIdx.y;
  for (int i = threadIdx.x; i < Cin * 7 * Wp; i += 256) {
    const int wp = i % Wp;
    const int r = i / Wp;               // c * 7 + ky
    const int c = r / 7, ky = r - c * 7;
    const int h = 2 * ho + ky - 3, w = wp - 3;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(in + ((static_cast<size_t>(n) * Cin + c) * H + h) * W + w);
    rows_s[i] = __float2bfloat16(v);
  }
  __syncthreads();
  const int nslots = 256 / K8;
  const int k8 = threadIdx.x % K8, slot = threadIdx.x / K8;
  if (slot >= nslots) return;
  int off[8];                           // smem offset of k at wo = 0 (column 2*wo + kx), or -1 for the zero padding of K
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k8 * 8 + j;
    if (k < K) {
      const int c = k / 49, r = k - c * 49;
      const int ky = r / 7, kx = r - ky * 7;
      off[j] = (c * 7 + ky) * Wp + kx;
    } else {
      off[j] = -1;
    }
  }
  __nv_bfloat16* orow = out + (static_cast<size_t>(n) * Ho + ho) * Wo * Kpad + k8 * 8;
  const unsigned short* rs = reinterpret_cast<const unsigned short*>(rows_s);
  for (int wo = slot; wo < Wo; wo += nslots) {
    unsigned short v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? rs[off[j] + 2 * wo] : static_cast<unsigned short>(0);
    uint4 u;
    u.x = v[0] | (static_cast<uint32_t>(v[1]) << 16);
    u.y = v[2] | (static_cast<uint32_t>(v[3]) << 16);
    u.z = v[4] | (static_cast<uint32_t>(v[5]) << 16);
    u.w = v[6] | (static_cast<uint32_t>(v[7]) << 16);
    *reinterpret_cast<uint4*>(orow + static_cast<size_t>(wo) * Kpad) = u;
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(3, 2, 1) (MT/model/backbone/resnet.py:68), NHWC; argmax position (0..8) kept for backward.
// ------------------------------------------------------------------------------------------------
// grid = (ceil(Wo*C8 / 256), Ho, N): row / image come from the block index, everything else is 32-bit arithmetic
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                          __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                                                          int N, int H, int W, int C, int Ho, int Wo) {
  const int C8 = C >> 3;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= Wo * C8) return;
  const int wo = t / C8, c8 = t - wo * C8;
  const int ho = blockIdx.y, n = blockIdx.z;
  const __nv_bfloat16* xin = x + static_cast<size_t>(n) * H * W * C + c8 * 8;
  uint4 raw[3][3];
  bool ok[3][3];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int h = 2 * ho + ky - 1, w = 2 * wo + kx - 1;
      ok[ky][kx] = h >= 0 && h < H && w >= 0 && w < W;
      if (ok[ky][kx]) raw[ky][kx] = ldg16(xin + (h * W + w) * C);
    }
  float best[8];
  int bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      if (!ok[ky][kx]) continue;
      float v[8];
      cvt8(raw[ky][kx], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (v[j] > best[j]) { best[j] = v[j]; bi[j] = ky * 3 + kx; }   // first maximum wins (ATen order)
      }
    }
  const size_t o = ((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * C + c8 * 8;
  store8(y + o, best);
  uint2 packed;
  packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
  packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
  *reinterpret_cast<uint2*>(idx + o) = packed;
}

// gather form: every input element sums the dy of the (<= 4) windows that selected it.  grid = (ceil(W*C8/256), H, N)
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                          const uint8_t* __restrict__ idx,
                                                          __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C,
                                                          int Ho, int Wo) {
  const int C8 = C >> 3;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= W * C8) return;
  const int w = t / C8, c8 = t - w * C8;
  const int h = blockIdx.y, n = blockIdx.z;
  // windows (ho, wo) with 2*ho-1 <= h <= 2*ho+1: ho in {h>>1, (h+1)>>1} (equal for even h), same for columns
  const int ho0 = h >> 1, ho1 = (h + 1) >> 1, wo0 = w >> 1, wo1 = (w + 1) >> 1;
  const size_t img = static_cast<size_t>(n) * Ho * Wo * C + c8 * 8;
  uint2 pk[2][2];
  uint4 gv[2][2];
  bool ok[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ho = a ? ho1 : ho0, wo = b ? wo1 : wo0;
      ok[a][b] = ho < Ho && wo < Wo && (a == 0 || ho1 != ho0) && (b == 0 || wo1 != wo0);
      if (ok[a][b]) {
        const size_t o = img + (static_cast<size_t>(ho) * Wo + wo) * C;
        pk[a][b] = __ldg(reinterpret_cast<const uint2*>(idx + o));
        gv[a][b] = ldg16(dy + o);
      }
    }
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      if (!ok[a][b]) continue;
      const int ho = a ? ho1 : ho0, wo = b ? wo1 : wo0;
      const int code = (h - 2 * ho + 1) * 3 + (w - 2 * wo + 1);
      float g[8];
      cvt8(gv[a][b], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sel = ((j < 4 ? pk[a][b].x : pk[a][b].y) >> (8 * (j & 3))) & 0xff;
        if (sel == code) acc[j] += g[j];
      }
    }
  store8(dx + ((static_cast<size_t>(n) * H + h) * W + w) * C + c8 * 8, acc);
}

// ------------------------------------------------------------------------------------------------
// SE fusion (MT/model/utils.py:84-95, MT/model/encoder_fusion.py:63-90): squeeze MLP and weighted add.
// ------------------------------------------------------------------------------------------------
// gap[n][c] += sum over pixels x   (used where the squeeze is not fused into bn_apply)
__global__ void __launch_bounds__(256) gap_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ gap, int HW,
                                                  int C, int chunks) {
  extern __shared__ float red[];
  const int C8 = C >> 3;
  const int planes = 256 / C8;
  const int c8 = threadIdx.x % C8, plane = threadIdx.x / C8;
  const int n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  for (int p = p0 + plane; p < p1 && plane < planes; p += planes) {
    float v[8];
    load8(x + (static_cast<size_t>(n) * HW + p) * C + c8 * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += v[j];
  }
  block_channel_reduce<1>(acc, c8, C8, plane, planes, red);
  if (plane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(gap + static_cast<size_t>(n) * C + c8 * 8 + j, red[c8 * 8 + j]);
  }
}

// one block per image: mean = gap/HW ; hid = relu(W1 mean + b1) ; w = sigmoid(W2 hid + b2).  gap is zeroed.
__global__ void se_mlp_fwd_kernel(float* __restrict__ gap, float inv_hw, const float* __restrict__ w1,
                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                  const float* __restrict__ b2, float* __restrict__ mean_out,
                                  float* __restrict__ hid_out, float* __restrict__ wgt_out, int C, int Cr) {
  extern __shared__ float sm[];   // mean[C] | hid[Cr]
  float* mean = sm;
  float* hid = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m = gap[static_cast<size_t>(n) * C + c] * inv_hw;
    gap[static_cast<size_t>(n) * C + c] = 0.f;
    mean[c] = m;
    mean_out[static_cast<size_t>(n) * C + c] = m;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < Cr; r += nwarps) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += w1[r * C + c] * mean[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s = fmaxf(s + b1[r], 0.f);
      hid[r] = s;
      hid_out[static_cast<size_t>(n) * Cr + r] = s;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = b2[c];
    for (int r = 0; r < Cr; ++r) s += w2[c * Cr + r] * hid[r];
    wgt_out[static_cast<size_t>(n) * C + c] = 1.f / (1.f + expf(-s));
  }
}

// backward of the squeeze MLP for one image per block; parameter gradients via atomics.
// dwgt is zeroed after use; dmean_out = gradient wrt the per-pixel mean, already divided by HW.
__global__ void se_mlp_bwd_kernel(float* __restrict__ dwgt, const float* __restrict__ wgt,
                                  const float* __restrict__ hid, const float* __restrict__ mean, float inv_hw,
                                  const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ dw1,
                                  float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2,
                                  float* __restrict__ dmean_out, int C, int Cr) {
  extern __shared__ float sm[];   // dz2[C] | dhid[Cr]
  float* dz2 = sm;
  float* dh = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const size_t i = static_cast<size_t>(n) * C + c;
    const float s = wgt[i];
    const float d = dwgt[i] * s * (1.f - s);
    dwgt[i] = 0.f;
    dz2[c] = d;
    atomicAdd(db2 + c, d);
    for (int r = 0; r < Cr; ++r) atomicAdd(dw2 + c * Cr + r, d * hid[static_cast<size_t>(n) * Cr + r]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < Cr; r += nwarps) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += w2[c * Cr + r] * dz2[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      s = hid[static_cast<size_t>(n) * Cr + r] > 0.f ? s : 0.f;
      dh[r] = s;
      atomicAdd(db1 + r, s);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    const float m = mean[static_cast<size_t>(n) * C + c];
    for (int r = 0; r < Cr; ++r) {
      s += w1[r * C + c] * dh[r];
      atomicAdd(dw1 + r * C + c, dh[r] * m);
    }
    dmean_out[static_cast<size_t>(n) * C + c] = s * inv_hw;
  }
}

// out = a * wa[n,c] + b * wb[n,c]
__global__ void __launch_bounds__(256) se_fuse_fwd_kernel(const __nv_bfloat16* __restrict__ a,
                                                          const __nv_bfloat16* __restrict__ b,
                                                          const float* __restrict__ wa, const float* __restrict__ wb,
                                                          __nv_bfloat16* __restrict__ out, int N, int HW, int C) {
  const int C8 = C >> 3;
  const size_t total = static_cast<size_t>(N) * HW * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    const int n = static_cast<int>(i / C8 / HW);
    float va[8], vb[8], fa[8], fb[8], o[8];
    load8(a + i * 8, va);
    load8(b + i * 8, vb);
    load8f(wa + static_cast<size_t>(n) * C + c8 * 8, fa);
    load8f(wb + static_cast<size_t>(n) * C + c8 * 8, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = va[j] * fa[j] + vb[j] * fb[j];
    store8(out + i * 8, o);
  }
}

// dwa[n,c] += sum_p dout*a ; dwb[n,c] += sum_p dout*b
__global__ void __launch_bounds__(256) se_fuse_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                 const __nv_bfloat16* __restrict__ a,
                                                                 const __nv_bfloat16* __restrict__ b,
                                                                 float* __restrict__ dwa, float* __restrict__ dwb,
                                                                 int HW, int C, int chunks) {
  extern __shared__ float red[];
  const int C8 = C >> 3;
  const int planes = 256 / C8;
  const int c8 = threadIdx.x % C8, plane = threadIdx.x / C8;
  const int n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  for (int p = p0 + plane; p < p1 && plane < planes; p += planes) {
    const size_t o = (static_cast<size_t>(n) * HW + p) * C + c8 * 8;
    float g[8], va[8], vb[8];
    load8(dout + o, g);
    load8(a + o, va);
    load8(b + o, vb);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] += g[j] * va[j]; acc[1][j] += g[j] * vb[j]; }
  }
  block_channel_reduce<2>(acc, c8, C8, plane, planes, red);
  if (plane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(dwa + static_cast<size_t>(n) * C + c8 * 8 + j, red[c8 * 8 + j]);
      atomicAdd(dwb + static_cast<size_t>(n) * C + c8 * 8 + j, red[C + c8 * 8 + j]);
    }
  }
}

// da = dout*wa + dmean_a ; db = dout*wb + dmean_b (+ db_prev)
__global__ void __launch_bounds__(256) se_fuse_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ wa, const float* __restrict__ wb,
    const float* __restrict__ dmean_a, const float* __restrict__ dmean_b, const __nv_bfloat16* __restrict__ db_prev,
    __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ db, int N, int HW, int C) {
  const int C8 = C >> 3;
  const size_t total = static_cast<size_t>(N) * HW * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    const int n = static_cast<int>(i / C8 / HW);
    const size_t nc = static_cast<size_t>(n) * C + c8 * 8;
    float g[8], fa[8], fb[8], ma[8], mb[8], oa[8], ob[8];
    load8(dout + i * 8, g);
    load8f(wa + nc, fa);
    load8f(wb + nc, fb);
    load8f(dmean_a + nc, ma);
    load8f(dmean_b + nc, mb);
#pragma unroll
    for (int j = 0; j < 8; ++j) { oa[j] = g[j] * fa[j] + ma[j]; ob[j] = g[j] * fb[j] + mb[j]; }
    if (db_prev) {
      float pv[8];
      load8(db_prev + i * 8, pv);
#pragma unroll
      for (int j = 0; j < 8; ++j) ob[j] += pv[j];
    }
    store8(da + i * 8, oa);
    store8(db + i * 8, ob);
  }
}

// ------------------------------------------------------------------------------------------------
// Pyramid pooling (MT/model/context_module/ppm.py:57-78): adaptive average pooling and bilinear
// (align_corners=False) upsampling, NHWC.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ap_start(int i, int in, int out) { return (i * in) / out; }
__device__ __forceinline__ int ap_end(int i, int in, int out) { return ((i + 1) * in + out - 1) / out; }

__global__ void __launch_bounds__(256) adaptive_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                __nv_bfloat16* __restrict__ y, int N, int H, int W,
                                                                int C, int B) {
  const int C8 = C >> 3;
  const size_t total = static_cast<size_t>(N) * B * B * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    size_t cell = i / C8;
    const int bx = static_cast<int>(cell % B);
    cell /= B;
    const int by = static_cast<int>(cell % B);
    const int n = static_cast<int>(cell / B);
    const int h0 = ap_start(by, H, B), h1 = ap_end(by, H, B), w0 = ap_start(bx, W, B), w1 = ap_end(bx, W, B);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        float v[8];
        load8(x + ((static_cast<size_t>(n) * H + h) * W + w) * C + c8 * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    const float inv = 1.f / static_cast<float>((h1 - h0) * (w1 - w0));
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    store8(y + i * 8, acc);
  }
}

// warp-per-cell variant for large windows (PPM bins over a 15x20 map: up to 300 pixels per cell, few cells): the 32
// lanes split the window, one shuffle reduction at the end
__global__ void __launch_bounds__(256) adaptive_pool_fwd_warp_kernel(const __nv_bfloat16* __restrict__ x,
                                                                     __nv_bfloat16* __restrict__ y, int N, int H, int W,
                                                                     int C, int B) {
  const int C8 = C >> 3;
  const int lane = threadIdx.x & 31;
  const long long item = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  if (item >= static_cast<long long>(N) * B * B * C8) return;
  const int c8 = static_cast<int>(item % C8);
  long long cell = item / C8;
  const int bx = static_cast<int>(cell % B);
  cell /= B;
  const int by = static_cast<int>(cell % B);
  const int n = static_cast<int>(cell / B);
  const int h0 = ap_start(by, H, B), h1 = ap_end(by, H, B), w0 = ap_start(bx, W, B), w1 = ap_end(bx, W, B);
  const int ww = w1 - w0, cnt = (h1 - h0) * ww;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int q = lane; q < cnt; q += 32) {
    const int h = h0 + q / ww, w = w0 + q % ww;
    float v[8];
    load8(x + ((static_cast<size_t>(n) * H + h) * W + w) * C + c8 * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
  if (lane == 0) {
    const float inv = 1.f / static_cast<float>(cnt);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    store8(y + item * 8, acc);
  }
}

// dx[n,h,w,c] (+)= sum over cells containing (h,w) of dy[cell]/area
__global__ void __launch_bounds__(256) adaptive_pool_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                __nv_bfloat16* __restrict__ dx, int N, int H, int W,
                                                                int C, int B, int accumulate) {
  const int C8 = C >> 3;
  const size_t total = static_cast<size_t>(N) * H * W * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    size_t pix = i / C8;
    const int w = static_cast<int>(pix % W);
    pix /= W;
    const int h = static_cast<int>(pix % H);
    const int n = static_cast<int>(pix / H);
    float acc[8];
    if (accumulate) load8(dx + i * 8, acc);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
    for (int by = 0; by < B; ++by) {
      const int h0 = ap_start(by, H, B), h1 = ap_end(by, H, B);
      if (h < h0 || h >= h1) continue;
      for (int bx = 0; bx < B; ++bx) {
        const int w0 = ap_start(bx, W, B), w1 = ap_end(bx, W, B);
        if (w < w0 || w >= w1) continue;
        float g[8];
        load8(dy + ((static_cast<size_t>(n) * B + by) * B + bx) * C + c8 * 8, g);
        const float inv = 1.f / static_cast<float>((h1 - h0) * (w1 - w0));
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j] * inv;
      }
    }
    store8(dx + i * 8, acc);
  }
}

__device__ __forceinline__ void bilinear_src(int dst, int in, int out, int* i0, int* i1, float* lam) {
  float src = (static_cast<float>(dst) + 0.5f) * (static_cast<float>(in) / static_cast<float>(out)) - 0.5f;
  if (src < 0.f) src = 0.f;
  const int a = static_cast<int>(src);
  *i0 = a;
  *i1 = a + 1 < in ? a + 1 : in - 1;
  *lam = src - static_cast<float>(a);
}

// y[n,h,w, coff + c] = bilinear(x[n,:,:,c])   (y has channel pitch y_cs: writes straight into the concat buffer)
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ y, int N, int Hi, int Wi,
                                                           int Ho, int Wo, int C, int y_cs, int y_coff) {
  const int C8 = C >> 3;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    size_t pix = i / C8;
    const int w = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int h = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_src(h, Hi, Ho, &h0, &h1, &lh);
    bilinear_src(w, Wi, Wo, &w0, &w1, &lw);
    float v00[8], v01[8], v10[8], v11[8], o[8];
    const size_t base = static_cast<size_t>(n) * Hi * Wi;
    load8(x + (base + h0 * Wi + w0) * C + c8 * 8, v00);
    load8(x + (base + h0 * Wi + w1) * C + c8 * 8, v01);
    load8(x + (base + h1 * Wi + w0) * C + c8 * 8, v10);
    load8(x + (base + h1 * Wi + w1) * C + c8 * 8, v11);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.f - lh) * ((1.f - lw) * v00[j] + lw * v01[j]) + lh * ((1.f - lw) * v10[j] + lw * v11[j]);
    store8(y + ((static_cast<size_t>(n) * Ho + h) * Wo + w) * y_cs + y_coff + c8 * 8, o);
  }
}

// dx[n,hi,wi,c] = sum over destination pixels of weight * dy[n,h,w,coff+c]  (gather over all dst pixels; maps are tiny)
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           __nv_bfloat16* __restrict__ dx, int N, int Hi, int Wi,
                                                           int Ho, int Wo, int C, int dy_cs, int dy_coff) {
  const int C8 = C >> 3;
  const size_t total = static_cast<size_t>(N) * Hi * Wi * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    size_t pix = i / C8;
    const int wi = static_cast<int>(pix % Wi);
    pix /= Wi;
    const int hi = static_cast<int>(pix % Hi);
    const int n = static_cast<int>(pix / Hi);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int h = 0; h < Ho; ++h) {
      int h0, h1; float lh;
      bilinear_src(h, Hi, Ho, &h0, &h1, &lh);
      const float wh = (h0 == hi ? 1.f - lh : 0.f) + (h1 == hi ? lh : 0.f);
      if (wh == 0.f) continue;
      for (int w = 0; w < Wo; ++w) {
        int w0, w1; float lw;
        bilinear_src(w, Wi, Wo, &w0, &w1, &lw);
        const float ww = (w0 == wi ? 1.f - lw : 0.f) + (w1 == wi ? lw : 0.f);
        if (ww == 0.f) continue;
        float g[8];
        load8(dy + ((static_cast<size_t>(n) * Ho + h) * Wo + w) * dy_cs + dy_coff + c8 * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wh * ww * g[j];
      }
    }
    store8(dx + i * 8, acc);
  }
}

// warp-per-source-pixel variant (tiny source maps: the 1x1 / 5x5 PPM features receive gradient from all 300 pixels)
__global__ void __launch_bounds__(256) bilinear_bwd_warp_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                __nv_bfloat16* __restrict__ dx, int N, int Hi, int Wi,
                                                                int Ho, int Wo, int C, int dy_cs, int dy_coff) {
  const int C8 = C >> 3;
  const int lane = threadIdx.x & 31;
  const long long item = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  if (item >= static_cast<long long>(N) * Hi * Wi * C8) return;
  const int c8 = static_cast<int>(item % C8);
  long long pix = item / C8;
  const int wi = static_cast<int>(pix % Wi);
  pix /= Wi;
  const int hi = static_cast<int>(pix % Hi);
  const int n = static_cast<int>(pix / Hi);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int q = lane; q < Ho * Wo; q += 32) {
    const int h = q / Wo, w = q - h * Wo;
    int h0, h1, w0, w1; float lh, lw;
    bilinear_src(h, Hi, Ho, &h0, &h1, &lh);
    bilinear_src(w, Wi, Wo, &w0, &w1, &lw);
    const float wh = (h0 == hi ? 1.f - lh : 0.f) + (h1 == hi ? lh : 0.f);
    const float ww = (w0 == wi ? 1.f - lw : 0.f) + (w1 == wi ? lw : 0.f);
    if (wh == 0.f || ww == 0.f) continue;
    float g[8];
    load8(dy + ((static_cast<size_t>(n) * Ho + h) * Wo + w) * dy_cs + dy_coff + c8 * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += wh * ww * g[j];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
  if (lane == 0) store8(dx + item * 8, acc);
}

// (learned upsampling: upsample.cu)

// ------------------------------------------------------------------------------------------------
// Output boundary: NHWC bf16 -> NCHW fp32 (the reference's output convention), with the instance-head
// activations (MT/model/decoder/instance.py:113-119; MT/utils/_torch.py:88-91) fused; and the reverse for
// the incoming output gradients.
// act_mode 0: copy Creal channels.  act_mode 1: instance head — ch0 sigmoid, ch1-2 tanh, ch3-4 unit length.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y0,
                                                           float* __restrict__ y1, float* __restrict__ y2, int N,
                                                           int HW, int C, int Creal, int act_mode) {
  const size_t total = static_cast<size_t>(N) * HW;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = pix / HW, p = pix - n * HW;
    if (act_mode == 0) {
      for (int c8 = 0; c8 * 8 < Creal; ++c8) {
        float v[8];
        load8(x + pix * C + c8 * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c8 * 8 + j;
          if (c < Creal) y0[(n * Creal + c) * HW + p] = v[j];
        }
      }
    } else {
      float v[8];
      load8(x + pix * C, v);
      y0[n * HW + p] = 1.f / (1.f + expf(-v[0]));
      y1[(n * 2 + 0) * HW + p] = tanhf(v[1]);
      y1[(n * 2 + 1) * HW + p] = tanhf(v[2]);
      if (y2) {
        const float r = sqrtf(v[3] * v[3] + v[4] * v[4]) + 1e-7f;
        y2[(n * 2 + 0) * HW + p] = v[3] / r;
        y2[(n * 2 + 1) * HW + p] = v[4] / r;
      }
    }
  }
}

// act_mode 0 through shared memory: a block converts 256 consecutive pixels of one image.  NHWC side: the 256 x C bf16
// block is one contiguous run (16-byte vectors, fully coalesced); NCHW side: per channel 256 consecutive fp32.
// smem tile [256][C + 2] bf16 (odd word pitch: conflict-free column reads).
__global__ void __launch_bounds__(256) nhwc_to_nchw_tile_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                                                int HW, int C, int Creal) {
  extern __shared__ __nv_bfloat16 tile_s[];
  const int pitch = C + 2;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * 256;
  const int np = min(256, HW - p0);
  const int C8 = C >> 3;
  const __nv_bfloat16* src = x + (static_cast<size_t>(n) * HW + p0) * C;
  for (int v = threadIdx.x; v < np * C8; v += 256) {
    const int pix = v / C8, c8 = v - pix * C8;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + v);
    uint32_t* d = reinterpret_cast<uint32_t*>(tile_s + pix * pitch + c8 * 8);
    d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
  }
  __syncthreads();
  const int p = threadIdx.x;
  if (p < np) {
    float* dst = y + static_cast<size_t>(n) * Creal * HW + p0 + p;
    for (int c = 0; c < Creal; ++c) dst[static_cast<size_t>(c) * HW] = __bfloat162float(tile_s[p * pitch + c]);
  }
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_grad_tile_kernel(const float* __restrict__ g,
                                                                     __nv_bfloat16* __restrict__ dx, int HW, int C,
                                                                     int Creal) {
  extern __shared__ __nv_bfloat16 tile_s[];
  const int pitch = C + 2;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * 256;
  const int np = min(256, HW - p0);
  const int C8 = C >> 3;
  const int p = threadIdx.x;
  if (p < np) {
    const float* src = g + static_cast<size_t>(n) * Creal * HW + p0 + p;
    for (int c = 0; c < Creal; ++c) tile_s[p * pitch + c] = __float2bfloat16(__ldg(src + static_cast<size_t>(c) * HW));
    for (int c = Creal; c < C; ++c) tile_s[p * pitch + c] = __float2bfloat16(0.f);
  }
  __syncthreads();
  __nv_bfloat16* dst = dx + (static_cast<size_t>(n) * HW + p0) * C;
  for (int v = threadIdx.x; v < np * C8; v += 256) {
    const int pix = v / C8, c8 = v - pix * C8;
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(tile_s + pix * pitch + c8 * 8);
    reinterpret_cast<uint4*>(dst)[v] = make_uint4(sp[0], sp[1], sp[2], sp[3]);
  }
}

// gradient of the above: g* are NCHW fp32 output gradients (null = zero); x is the saved pre-activation map
__global__ void __launch_bounds__(256) nchw_to_nhwc_grad_kernel(const float* __restrict__ g0,
                                                                const float* __restrict__ g1,
                                                                const float* __restrict__ g2,
                                                                const __nv_bfloat16* __restrict__ x,
                                                                __nv_bfloat16* __restrict__ dx, int N, int HW, int C,
                                                                int Creal, int act_mode) {
  const size_t total = static_cast<size_t>(N) * HW;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = pix / HW, p = pix - n * HW;
    if (act_mode == 0) {
      for (int c8 = 0; c8 * 8 < C; ++c8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c8 * 8 + j;
          v[j] = (c < Creal && g0) ? __ldg(g0 + (n * Creal + c) * HW + p) : 0.f;
        }
        store8(dx + pix * C + c8 * 8, v);
      }
    } else {
      float v[8], o[8];
      load8(x + pix * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
      if (g0) {
        const float s = 1.f / (1.f + expf(-v[0]));
        o[0] = __ldg(g0 + n * HW + p) * s * (1.f - s);
      }
      if (g1) {
        const float t1 = tanhf(v[1]), t2 = tanhf(v[2]);
        o[1] = __ldg(g1 + (n * 2 + 0) * HW + p) * (1.f - t1 * t1);
        o[2] = __ldg(g1 + (n * 2 + 1) * HW + p) * (1.f - t2 * t2);
      }
      if (g2) {
        const float ga = __ldg(g2 + (n * 2 + 0) * HW + p), gb = __ldg(g2 + (n * 2 + 1) * HW + p);
        const float r = sqrtf(v[3] * v[3] + v[4] * v[4]);
        const float re = r + 1e-7f;
        const float dot = ga * v[3] + gb * v[4];
        const float k = r > 0.f ? dot / (r * re * re) : 0.f;
        o[3] = ga / re - k * v[3];
        o[4] = gb / re - k * v[4];
      }
      store8(dx + pix * C, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Scene head: nn.Linear(256 -> n_classes) on the PPM bin-1 feature (MT/model/decoder/scene.py:32-65)
// ------------------------------------------------------------------------------------------------
__global__ void linear_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ y, int N, int K, int M) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N * M) return;
  const int n = warp / M, m = warp - n * M;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += __bfloat162float(x[static_cast<size_t>(n) * K + k]) * w[m * K + k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[n * M + m] = s + b[m];
}
// dx[n][k] = sum_m dy[n][m] w[m][k] ; dw[m][k] += sum_n dy[n][m] x[n][k] ; db[m] += sum_n dy[n][m]
__global__ void linear_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                  const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
                                  float* __restrict__ db, int N, int K, int M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * K) {
    const int n = i / K, k = i - n * K;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += dy[n * M + m] * w[m * K + k];
    dx[i] = __float2bfloat16(s);
  }
  if (i < M * K) {
    const int m = i / K, k = i - m * K;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dy[n * M + m] * __bfloat162float(x[static_cast<size_t>(n) * K + k]);
    dw[i] += s;
  }
  if (i < M) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dy[n * M + i];
    db[i] += s;
  }
}

// a += b (bf16, gradient fan-in); optional channel slice copy
__global__ void __launch_bounds__(256) add_inplace_kernel(__nv_bfloat16* __restrict__ a,
                                                          const __nv_bfloat16* __restrict__ b, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    load8(a + i * 8, x);
    load8(b + i * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    store8(a + i * 8, x);
  }
}
// dst[p][dcoff + c] (=|+=) src[p][scoff + c], c < C
__global__ void __launch_bounds__(256) copy_channels_kernel(const __nv_bfloat16* __restrict__ src,
                                                            __nv_bfloat16* __restrict__ dst, long long P, int C,
                                                            int scs, int scoff, int dcs, int dcoff, int accumulate) {
  const int C8 = C >> 3;
  const size_t total = static_cast<size_t>(P) * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    const size_t p = i / C8;
    float v[8];
    load8(src + p * scs + scoff + c8 * 8, v);
    if (accumulate) {
      float o[8];
      load8(dst + p * dcs + dcoff + c8 * 8, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += o[j];
    }
    store8(dst + p * dcs + dcoff + c8 * 8, v);
  }
}

}  // namespace eb

using namespace eb;
#define STREAM static_cast<cudaStream_t>(stream)

static inline int pick_chunks(int N, int HW, int planes) {
  // blocks per image so that the grid is ~4 waves and every block still has >= 8 pixel iterations
  static int waves = -1;
  if (waves < 0) { const char* e = getenv("EB200_BN_WAVES"); waves = e ? atoi(e) : 8; if (waves < 1) waves = 8; }
  int chunks = (waves * num_sms() + N - 1) / N;
  const int maxc = (HW + planes * 8 - 1) / (planes * 8);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  return chunks;
}

extern "C" int eb200_bn_finalize(float* stats, long long count, const float* gamma, const float* beta, float eps,
                                 float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                                 float* mean, float* rstd, int C, void* stream) {
  EB_REQUIRE(stats && gamma && beta && scale && shift && mean && rstd && C > 0, "eb200_bn_finalize: bad argument");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, STREAM>>>(stats, static_cast<float>(count), gamma, beta, eps, momentum,
                                                           running_mean, running_var, scale, shift, mean, rstd, C);
  return launch_check("bn_finalize_kernel");
}

extern "C" int eb200_bn_apply(const void* x, void* y, const float* scale, const float* shift, const float* drop,
                              const void* res_pre, const void* res_post, float* gap, int N, int HW, int C, int y_cs,
                              int y_coff, int relu, void* stream) {
  EB_REQUIRE(x && y && scale && shift, "eb200_bn_apply: null argument");
  EB_REQUIRE(C % 8 == 0 && C <= 2048, "eb200_bn_apply: C=%d unsupported", C);
  BnApplyArgs a;
  memset(&a, 0, sizeof(a));
  a.x = static_cast<const __nv_bfloat16*>(x); a.y = static_cast<__nv_bfloat16*>(y);
  a.scale = scale; a.shift = shift; a.drop = drop;
  a.res_pre = static_cast<const __nv_bfloat16*>(res_pre); a.res_post = static_cast<const __nv_bfloat16*>(res_post);
  a.gap = gap; a.N = N; a.HW = HW; a.C = C; a.y_cs = y_cs > 0 ? y_cs : C; a.y_coff = y_coff; a.relu = relu;
  const int planes = 256 / (C / 8);
  a.chunks = pick_chunks(N, HW, planes);
  const size_t smem = gap ? static_cast<size_t>(planes) * C * sizeof(float) : 0;
  {
    void* kargs[1] = {&a};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_apply_kernel), dim3(N * a.chunks), dim3(256), smem, STREAM, kargs, 1, 8));
  }
  return launch_check("bn_apply_kernel");
}

extern "C" int eb200_bn_apply_train(const void* x, void* y, const float* stats, long long count, const float* gamma,
                                    const float* beta, float eps, float momentum, float* running_mean,
                                    float* running_var, float* scale_out, float* shift_out, float* mean_out,
                                    float* rstd_out, const float* drop, const void* res_pre, const void* res_post,
                                    float* gap, int N, int HW, int C, int y_cs, int y_coff, int relu, void* stream) {
  EB_REQUIRE(x && y && stats && gamma && beta && scale_out && shift_out && mean_out && rstd_out,
             "eb200_bn_apply_train: null argument");
  EB_REQUIRE(C % 8 == 0 && C <= 2048, "eb200_bn_apply_train: C=%d unsupported", C);
  BnApplyArgs a;
  memset(&a, 0, sizeof(a));
  a.x = static_cast<const __nv_bfloat16*>(x); a.y = static_cast<__nv_bfloat16*>(y);
  a.drop = drop;
  a.res_pre = static_cast<const __nv_bfloat16*>(res_pre); a.res_post = static_cast<const __nv_bfloat16*>(res_post);
  a.gap = gap; a.N = N; a.HW = HW; a.C = C; a.y_cs = y_cs > 0 ? y_cs : C; a.y_coff = y_coff; a.relu = relu;
  a.stats = stats; a.gamma = gamma; a.beta = beta; a.running_mean = running_mean; a.running_var = running_var;
  a.scale_out = scale_out; a.shift_out = shift_out; a.mean_out = mean_out; a.rstd_out = rstd_out;
  a.count = static_cast<float>(count); a.eps = eps; a.momentum = momentum;
  const int planes = 256 / (C / 8);
  a.chunks = pick_chunks(N, HW, planes);
  const size_t smem = gap ? static_cast<size_t>(planes) * C * sizeof(float) : 0;
  {
    void* kargs[1] = {&a};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_apply_kernel), dim3(N * a.chunks), dim3(256), smem, STREAM, kargs, 1, 8));
  }
  return launch_check("bn_apply_kernel");
}

static int fill_bn_bwd(BnBwdArgs& a, const void* dy, const void* x, const void* mask_src, const float* drop,
                       const float* mean, const float* rstd, const float* scale, const float* shift,
                       const float* gamma, float* sums, int N, int HW, int C, int dy_cs, int dy_coff, int relu_mode) {
  EB_REQUIRE(dy && x && mean && rstd && scale && shift && sums, "bn backward: null argument");
  EB_REQUIRE(C % 8 == 0 && C <= 2048, "bn backward: C=%d unsupported", C);
  EB_REQUIRE(relu_mode != 1 || mask_src, "bn backward: relu_mode 1 needs mask_src");
  a.dy = static_cast<const __nv_bfloat16*>(dy); a.x = static_cast<const __nv_bfloat16*>(x);
  a.mask_src = static_cast<const __nv_bfloat16*>(mask_src); a.drop = drop; a.mean = mean; a.rstd = rstd;
  a.scale = scale; a.shift = shift; a.gamma = gamma; a.sums = sums; a.dx = nullptr; a.dres = nullptr;
  a.N = N; a.HW = HW; a.C = C; a.dy_cs = dy_cs > 0 ? dy_cs : C; a.dy_coff = dy_coff; a.relu_mode = relu_mode;
  a.chunks = pick_chunks(N, HW, 256 / (C / 8));
  a.inv_count = 1.f / (static_cast<float>(N) * static_cast<float>(HW));
  a.replicas = 0; a.dgamma = nullptr; a.dbeta = nullptr; a.raw_sums = 0; a.fused = 0;
  return 0;
}

static inline int pick_chunks_reduce(int N, int HW, int planes) {
  int chunks = (2 * num_sms()) / N;   // one wave at 2 resident blocks per SM
  const int maxc = (HW + planes * 8 - 1) / (planes * 8);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  return chunks;
}

extern "C" int eb200_bn_bwd_reduce(const void* dy, const void* x, const void* mask_src, const float* drop,
                                   const float* mean, const float* rstd, const float* scale, const float* shift,
                                   float* partials, long long partials_floats, float* sums, float* dgamma,
                                   float* dbeta, int N, int HW, int C, int dy_cs, int dy_coff, int relu_mode,
                                   void* stream) {
  BnBwdArgs a;
  EB_REQUIRE(partials && dgamma && dbeta, "eb200_bn_bwd_reduce: null argument");
  if (fill_bn_bwd(a, dy, x, mask_src, drop, mean, rstd, scale, shift, nullptr, partials, N, HW, C, dy_cs, dy_coff,
                  relu_mode))
    return 1;
  a.chunks = pick_chunks_reduce(N, HW, 256 / (C / 8));
  const int nblocks = N * a.chunks;
  EB_REQUIRE(static_cast<long long>(nblocks) * 2 * C <= partials_floats,
             "eb200_bn_bwd_reduce: partials workspace too small (%lld floats needed)",
             static_cast<long long>(nblocks) * 2 * C);
  const size_t smem = static_cast<size_t>(256 / (C / 8)) * 2 * C * sizeof(float);
  {
    void* kargs[1] = {&a};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_bwd_reduce_kernel), dim3(nblocks), dim3(256), smem, STREAM, kargs, 1, 8));
  }
  if (launch_check("bn_bwd_reduce_kernel")) return 1;
  bn_bwd_combine_kernel<<<ceil_div(C, 32), 256, 0, STREAM>>>(partials, nblocks, sums, dgamma, dbeta, mean, rstd, C);
  return launch_check("bn_bwd_combine_kernel");
}

extern "C" int eb200_bn_bwd_apply(const void* dy, const void* x, const void* mask_src, const float* drop,
                                  const float* mean, const float* rstd, const float* scale, const float* shift,
                                  const float* gamma, const float* sums, void* dx, void* dres, int N, int HW, int C,
                                  int dy_cs, int dy_coff, int relu_mode, void* stream) {
  BnBwdArgs a;
  EB_REQUIRE(gamma && dx, "eb200_bn_bwd_apply: null argument");
  if (fill_bn_bwd(a, dy, x, mask_src, drop, mean, rstd, scale, shift, gamma, const_cast<float*>(sums), N, HW, C, dy_cs,
                  dy_coff, relu_mode))
    return 1;
  a.dx = static_cast<__nv_bfloat16*>(dx);
  a.dres = static_cast<__nv_bfloat16*>(dres);
  {
    void* kargs[1] = {&a};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_bwd_apply_kernel), dim3(N * a.chunks), dim3(256), 0, STREAM, kargs, 1, 8));
  }
  return launch_check("bn_bwd_apply_kernel");
}

/* replica form: no partials workspace, no combine launch.  `ws` is fp32 [(replicas + 1) * 2C + 4], ZERO on entry:
 * replicas x [2C] block-sum targets | folded [2C] = (sum g, sum g*xhat) written by the last block | counter. */
extern "C" int eb200_bn_bwd_reduce_rep(const void* dy, const void* x, const void* mask_src, const float* drop,
                                       const float* mean, const float* rstd, const float* scale, const float* shift,
                                       float* ws, int replicas, float* dgamma, float* dbeta, int N, int HW, int C,
                                       int dy_cs, int dy_coff, int relu_mode, void* stream) {
  BnBwdArgs a;
  EB_REQUIRE(ws && replicas > 0 && dgamma && dbeta, "eb200_bn_bwd_reduce_rep: bad argument");
  if (fill_bn_bwd(a, dy, x, mask_src, drop, mean, rstd, scale, shift, nullptr, ws, N, HW, C, dy_cs, dy_coff, relu_mode))
    return 1;
  a.replicas = replicas; a.dgamma = dgamma; a.dbeta = dbeta;
  a.chunks = pick_chunks_reduce(N, HW, 256 / (C / 8));
  const size_t smem = static_cast<size_t>(256 / (C / 8)) * 2 * C * sizeof(float);
  {
    void* kargs[1] = {&a};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_bwd_reduce_kernel), dim3(N * a.chunks), dim3(256), smem, STREAM,
                      kargs, 1, 8));
  }
  return launch_check("bn_bwd_reduce_kernel");
}

/* reduce + apply in ONE launch (grid-wide barrier between the phases; falls back to two launches when the grid would not
 * be fully resident).  `ws` as for eb200_bn_bwd_reduce_rep, ZERO on entry. */
extern "C" int eb200_bn_bwd_fused(const void* dy, const void* x, const void* mask_src, const float* drop,
                                  const float* mean, const float* rstd, const float* scale, const float* shift,
                                  const float* gamma, float* ws, int replicas, float* dgamma, float* dbeta, void* dx,
                                  void* dres, int N, int HW, int C, int dy_cs, int dy_coff, int relu_mode,
                                  void* stream) {
  BnBwdArgs a;
  EB_REQUIRE(ws && replicas > 0 && dgamma && dbeta && gamma && dx, "eb200_bn_bwd_fused: bad argument");
  if (fill_bn_bwd(a, dy, x, mask_src, drop, mean, rstd, scale, shift, gamma, ws, N, HW, C, dy_cs, dy_coff, relu_mode))
    return 1;
  a.replicas = replicas; a.dgamma = dgamma; a.dbeta = dbeta;
  a.dx = static_cast<__nv_bfloat16*>(dx);
  a.dres = static_cast<__nv_bfloat16*>(dres);
  a.chunks = pick_chunks_reduce(N, HW, 256 / (C / 8));
  const size_t smem = static_cast<size_t>(256 / (C / 8)) * 2 * C * sizeof(float);
  const int nblocks = N * a.chunks;
  static int resident_per_sm = -1;
  int per_sm = 0;
  EB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_reduce_kernel, 256, smem));
  (void)resident_per_sm;
  // Only for small activations (<= 6 MB by default: the 15x20 maps; measured -0.2 ms per step, +-0 at 12 MB, +0.5 ms at 45 MB): there the kernels are latency bound and the second pass hits L1/L2;
  // on large tensors the 2-blocks-per-SM grid of the reduce under-feeds the bandwidth-bound apply phase (measured).
  static double max_mb = -1.0;
  if (max_mb < 0) { const char* e = getenv("EB200_BN_FUSED_MAX_MB"); max_mb = e ? atof(e) : 6.0; }
  const double mb = 2.0 * N * HW * C / 1e6;
  if (getenv("EB200_NO_BN_FUSED") || mb > max_mb || nblocks > per_sm * num_sms()) {
    // not fully resident: two launches
    {
      void* kargs[1] = {&a};
      EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_bwd_reduce_kernel), dim3(nblocks), dim3(256), smem, STREAM, kargs, 1, 8));
    }
    if (launch_check("bn_bwd_reduce_kernel")) return 1;
    BnBwdArgs b = a;
    b.replicas = 0;
    b.sums = ws + static_cast<size_t>(replicas) * 2 * C;
    b.chunks = pick_chunks(N, HW, 256 / (C / 8));
    void* kargs[1] = {&b};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_bwd_apply_kernel), dim3(N * b.chunks), dim3(256), 0, STREAM, kargs, 1, 8));
    return launch_check("bn_bwd_apply_kernel");
  }
  a.fused = 1;
  void* kargs[1] = {&a};
  EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_bwd_reduce_kernel), dim3(nblocks), dim3(256), smem, STREAM, kargs, 1, 8));
  return launch_check("bn_bwd_reduce_kernel(fused)");
}

extern "C" int eb200_bn_bwd_apply_raw(const void* g, const void* x, const float* mean, const float* rstd,
                                      const float* gamma, const float* raw_sums, float* dgamma, float* dbeta, void* dx,
                                      int N, int HW, int C, void* stream) {
  BnBwdArgs a;
  EB_REQUIRE(gamma && dx && dgamma && dbeta && raw_sums, "eb200_bn_bwd_apply_raw: null argument");
  // scale / shift are not read with relu_mode 0: pass mean / rstd as stand-ins for the null check
  if (fill_bn_bwd(a, g, x, nullptr, nullptr, mean, rstd, mean, rstd, gamma, const_cast<float*>(raw_sums), N, HW, C, 0, 0, 0))
    return 1;
  a.raw_sums = 1; a.dgamma = dgamma; a.dbeta = dbeta;
  a.dx = static_cast<__nv_bfloat16*>(dx);
  a.dres = nullptr;
  {
    void* kargs[1] = {&a};
    EB_CUDA(launch_ex(reinterpret_cast<const void*>(bn_bwd_apply_kernel), dim3(N * a.chunks), dim3(256), 0, STREAM, kargs, 1, 8));
  }
  return launch_check("bn_bwd_apply_kernel");
}

extern "C" int eb200_bn_bwd_param(float* sums, float* dgamma, float* dbeta, int C, void* stream) {
  EB_REQUIRE(sums && dgamma && dbeta, "eb200_bn_bwd_param: null argument");
  bn_bwd_param_kernel<<<ceil_div(C, 128), 128, 0, STREAM>>>(sums, dgamma, dbeta, C);
  return launch_check("bn_bwd_param_kernel");
}

extern "C" int eb200_colsum(const void* x, float* out, long long P, int C, int cs, int coff, void* stream) {
  EB_REQUIRE(x && out && C % 8 == 0 && C / 8 <= 256, "eb200_colsum: bad argument");
  const int C8 = C / 8;
  const int planes = 256 / C8;
  const int pp = planes;
  int grid = static_cast<int>((P + planes * 16 - 1) / (planes * 16));
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  if (grid < 1) grid = 1;
  colsum_kernel<<<grid, 256, static_cast<size_t>(pp) * C * sizeof(float), STREAM>>>(
      static_cast<const __nv_bfloat16*>(x), out, P, C, cs > 0 ? cs : C, coff);
  return launch_check("colsum_kernel");
}

extern "C" int eb200_im2col_stem(const float* in, void* out, int N, int Cin, int H, int W, int Kpad, void* stream) {
  EB_REQUIRE(in && out && Kpad % 8 == 0 && Kpad >= Cin * 49, "eb200_im2col_stem: bad argument");
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  EB_REQUIRE(Kpad / 8 <= 256 && N <= 65535, "eb200_im2col_stem: Kpad / N too large");
  const size_t smem = static_cast<size_t>(Cin) * 7 * (W + 6) * 2;
  EB_REQUIRE(smem <= 200 * 1024, "eb200_im2col_stem: input rows do not fit in shared memory (Cin=%d W=%d)", Cin, W);
  static bool configured = false;
  if (!configured) {
    EB_CUDA(cudaFuncSetAttribute(im2col_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  im2col_stem_kernel<<<dim3(Ho, N), 256, smem, STREAM>>>(in, static_cast<__nv_bfloat16*>(out), N, Cin, H, W, Ho, Wo,
                                                         Kpad);
  return launch_check("im2col_stem_kernel");
}

extern "C" int eb200_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, void* stream) {
  EB_REQUIRE(x && y && idx && C % 8 == 0, "eb200_maxpool_fwd: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  EB_REQUIRE(Ho <= 65535 && N <= 65535, "eb200_maxpool_fwd: extent too large");
  maxpool_fwd_kernel<<<dim3(ceil_div(Wo * (C / 8), 256), Ho, N), 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(x),
                                                                   static_cast<__nv_bfloat16*>(y),
                                                                   static_cast<uint8_t*>(idx), N, H, W, C, Ho, Wo);
  return launch_check("maxpool_fwd_kernel");
}
extern "C" int eb200_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, void* stream) {
  EB_REQUIRE(dy && dx && idx && C % 8 == 0, "eb200_maxpool_bwd: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  EB_REQUIRE(H <= 65535 && N <= 65535, "eb200_maxpool_bwd: extent too large");
  maxpool_bwd_kernel<<<dim3(ceil_div(W * (C / 8), 256), H, N), 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(dy),
                                                                   static_cast<const uint8_t*>(idx),
                                                                   static_cast<__nv_bfloat16*>(dx), N, H, W, C, Ho, Wo);
  return launch_check("maxpool_bwd_kernel");
}

extern "C" int eb200_gap(const void* x, float* gap, int N, int HW, int C, void* stream) {
  EB_REQUIRE(x && gap && C % 8 == 0 && C <= 2048, "eb200_gap: bad argument");
  const int planes = 256 / (C / 8);
  const int chunks = pick_chunks(N, HW, planes);
  gap_kernel<<<N * chunks, 256, static_cast<size_t>(planes) * C * sizeof(float), STREAM>>>(
      static_cast<const __nv_bfloat16*>(x), gap, HW, C, chunks);
  return launch_check("gap_kernel");
}
extern "C" int eb200_se_mlp_fwd(float* gap, int HW, const float* w1, const float* b1, const float* w2, const float* b2,
                                float* mean, float* hid, float* wgt, int N, int C, int Cr, void* stream) {
  EB_REQUIRE(gap && w1 && b1 && w2 && b2 && mean && hid && wgt, "eb200_se_mlp_fwd: null argument");
  se_mlp_fwd_kernel<<<N, 256, (C + Cr) * sizeof(float), STREAM>>>(gap, 1.f / HW, w1, b1, w2, b2, mean, hid, wgt, C, Cr);
  return launch_check("se_mlp_fwd_kernel");
}
extern "C" int eb200_se_mlp_bwd(float* dwgt, const float* wgt, const float* hid, const float* mean, int HW,
                                const float* w1, const float* w2, float* dw1, float* db1, float* dw2, float* db2,
                                float* dmean, int N, int C, int Cr, void* stream) {
  EB_REQUIRE(dwgt && wgt && hid && mean && w1 && w2 && dw1 && db1 && dw2 && db2 && dmean,
             "eb200_se_mlp_bwd: null argument");
  se_mlp_bwd_kernel<<<N, 256, (C + Cr) * sizeof(float), STREAM>>>(dwgt, wgt, hid, mean, 1.f / HW, w1, w2, dw1, db1, dw2,
                                                                  db2, dmean, C, Cr);
  return launch_check("se_mlp_bwd_kernel");
}
extern "C" int eb200_se_fuse_fwd(const void* a, const void* b, const float* wa, const float* wb, void* out, int N,
                                 int HW, int C, void* stream) {
  EB_REQUIRE(a && b && wa && wb && out && C % 8 == 0, "eb200_se_fuse_fwd: bad argument");
  const long long items = static_cast<long long>(N) * HW * (C / 8);
  se_fuse_fwd_kernel<<<grid_for(items, 256, 16), 256, 0, STREAM>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), wa, wb,
      static_cast<__nv_bfloat16*>(out), N, HW, C);
  return launch_check("se_fuse_fwd_kernel");
}
extern "C" int eb200_se_fuse_bwd_reduce(const void* dout, const void* a, const void* b, float* dwa, float* dwb, int N,
                                        int HW, int C, void* stream) {
  EB_REQUIRE(dout && a && b && dwa && dwb && C % 8 == 0 && C <= 2048, "eb200_se_fuse_bwd_reduce: bad argument");
  const int planes = 256 / (C / 8);
  const int chunks = pick_chunks(N, HW, planes);
  se_fuse_bwd_reduce_kernel<<<N * chunks, 256, static_cast<size_t>(planes) * 2 * C * sizeof(float), STREAM>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(a),
      static_cast<const __nv_bfloat16*>(b), dwa, dwb, HW, C, chunks);
  return launch_check("se_fuse_bwd_reduce_kernel");
}
extern "C" int eb200_se_fuse_bwd_apply(const void* dout, const float* wa, const float* wb, const float* dmean_a,
                                       const float* dmean_b, const void* db_prev, void* da, void* db, int N, int HW,
                                       int C, void* stream) {
  EB_REQUIRE(dout && wa && wb && dmean_a && dmean_b && da && db, "eb200_se_fuse_bwd_apply: null argument");
  const long long items = static_cast<long long>(N) * HW * (C / 8);
  se_fuse_bwd_apply_kernel<<<grid_for(items, 256, 16), 256, 0, STREAM>>>(
      static_cast<const __nv_bfloat16*>(dout), wa, wb, dmean_a, dmean_b, static_cast<const __nv_bfloat16*>(db_prev),
      static_cast<__nv_bfloat16*>(da), static_cast<__nv_bfloat16*>(db), N, HW, C);
  return launch_check("se_fuse_bwd_apply_kernel");
}

extern "C" int eb200_adaptive_pool_fwd(const void* x, void* y, int N, int H, int W, int C, int B, void* stream) {
  EB_REQUIRE(x && y && C % 8 == 0 && B >= 1, "eb200_adaptive_pool_fwd: bad argument");
  const long long items = static_cast<long long>(N) * B * B * (C / 8);
  if ((H / B) * (W / B) >= 8 && items * 32 < (1ll << 30)) {     // big windows: a warp per cell
    adaptive_pool_fwd_warp_kernel<<<static_cast<int>((items * 32 + 255) / 256), 256, 0, STREAM>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, H, W, C, B);
    return launch_check("adaptive_pool_fwd_warp_kernel");
  }
  adaptive_pool_fwd_kernel<<<grid_for(items, 256), 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(x),
                                                                     static_cast<__nv_bfloat16*>(y), N, H, W, C, B);
  return launch_check("adaptive_pool_fwd_kernel");
}
extern "C" int eb200_adaptive_pool_bwd(const void* dy, void* dx, int N, int H, int W, int C, int B, int accumulate,
                                       void* stream) {
  EB_REQUIRE(dy && dx && C % 8 == 0 && B >= 1, "eb200_adaptive_pool_bwd: bad argument");
  const long long items = static_cast<long long>(N) * H * W * (C / 8);
  adaptive_pool_bwd_kernel<<<grid_for(items, 256), 256, 0, STREAM>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, H, W, C, B, accumulate);
  return launch_check("adaptive_pool_bwd_kernel");
}
extern "C" int eb200_bilinear_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int y_cs,
                                  int y_coff, void* stream) {
  EB_REQUIRE(x && y && C % 8 == 0, "eb200_bilinear_fwd: bad argument");
  const long long items = static_cast<long long>(N) * Ho * Wo * (C / 8);
  bilinear_fwd_kernel<<<grid_for(items, 256), 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(x),
                                                                static_cast<__nv_bfloat16*>(y), N, Hi, Wi, Ho, Wo, C,
                                                                y_cs > 0 ? y_cs : C, y_coff);
  return launch_check("bilinear_fwd_kernel");
}
extern "C" int eb200_bilinear_bwd(const void* dy, void* dx, int N, int Hi, int Wi, int Ho, int Wo, int C, int dy_cs,
                                  int dy_coff, void* stream) {
  EB_REQUIRE(dy && dx && C % 8 == 0, "eb200_bilinear_bwd: bad argument");
  const long long items = static_cast<long long>(N) * Hi * Wi * (C / 8);
  if ((Ho / Hi) * (Wo / Wi) >= 8 && items * 32 < (1ll << 30)) {   // tiny source map: a warp per source pixel
    bilinear_bwd_warp_kernel<<<static_cast<int>((items * 32 + 255) / 256), 256, 0, STREAM>>>(
        static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, Hi, Wi, Ho, Wo, C,
        dy_cs > 0 ? dy_cs : C, dy_coff);
    return launch_check("bilinear_bwd_warp_kernel");
  }
  bilinear_bwd_kernel<<<grid_for(items, 256), 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(dy),
                                                                static_cast<__nv_bfloat16*>(dx), N, Hi, Wi, Ho, Wo, C,
                                                                dy_cs > 0 ? dy_cs : C, dy_coff);
  return launch_check("bilinear_bwd_kernel");
}

extern "C" int eb200_nhwc_to_nchw(const void* x, float* y0, float* y1, float* y2, int N, int HW, int C, int Creal,
                                  int act_mode, void* stream) {
  EB_REQUIRE(x && y0 && C % 8 == 0, "eb200_nhwc_to_nchw: bad argument");
  EB_REQUIRE(act_mode == 0 || (C == 8 && y1), "eb200_nhwc_to_nchw: instance mode needs C == 8");
  if (act_mode == 0 && C <= 256 && N <= 65535) {
    nhwc_to_nchw_tile_kernel<<<dim3(ceil_div(HW, 256), N), 256, static_cast<size_t>(256) * (C + 2) * 2, STREAM>>>(
        static_cast<const __nv_bfloat16*>(x), y0, HW, C, Creal);
    return launch_check("nhwc_to_nchw_tile_kernel");
  }
  nhwc_to_nchw_kernel<<<grid_for(static_cast<long long>(N) * HW, 256, 16), 256, 0, STREAM>>>(
      static_cast<const __nv_bfloat16*>(x), y0, y1, y2, N, HW, C, Creal, act_mode);
  return launch_check("nhwc_to_nchw_kernel");
}
extern "C" int eb200_nchw_to_nhwc_grad(const float* g0, const float* g1, const float* g2, const void* x, void* dx, int N,
                                       int HW, int C, int Creal, int act_mode, void* stream) {
  EB_REQUIRE(dx && C % 8 == 0, "eb200_nchw_to_nhwc_grad: bad argument");
  EB_REQUIRE(act_mode == 0 || (C == 8 && x), "eb200_nchw_to_nhwc_grad: instance mode needs C == 8 and x");
  if (act_mode == 0 && g0 && C <= 256 && N <= 65535) {
    nchw_to_nhwc_grad_tile_kernel<<<dim3(ceil_div(HW, 256), N), 256, static_cast<size_t>(256) * (C + 2) * 2, STREAM>>>(
        g0, static_cast<__nv_bfloat16*>(dx), HW, C, Creal);
    return launch_check("nchw_to_nhwc_grad_tile_kernel");
  }
  nchw_to_nhwc_grad_kernel<<<grid_for(static_cast<long long>(N) * HW, 256, 16), 256, 0, STREAM>>>(
      g0, g1, g2, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(dx), N, HW, C, Creal, act_mode);
  return launch_check("nchw_to_nhwc_grad_kernel");
}

extern "C" int eb200_linear_fwd(const void* x, const float* w, const float* b, float* y, int N, int K, int M,
                                void* stream) {
  EB_REQUIRE(x && w && b && y, "eb200_linear_fwd: null argument");
  linear_fwd_kernel<<<ceil_div(N * M * 32, 256), 256, 0, STREAM>>>(static_cast<const __nv_bfloat16*>(x), w, b, y, N, K, M);
  return launch_check("linear_fwd_kernel");
}
extern "C" int eb200_linear_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* db, int N,
                                int K, int M, void* stream) {
  EB_REQUIRE(dy && x && w && dx && dw && db, "eb200_linear_bwd: null argument");
  const int items = (N > M ? N : M) * K;
  linear_bwd_kernel<<<ceil_div(items, 256), 256, 0, STREAM>>>(dy, static_cast<const __nv_bfloat16*>(x), w,
                                                              static_cast<__nv_bfloat16*>(dx), dw, db, N, K, M);
  return launch_check("linear_bwd_kernel");
}

extern "C" int eb200_add_inplace(void* a, const void* b, long long n, void* stream) {
  EB_REQUIRE(a && b && n % 8 == 0, "eb200_add_inplace: bad argument");
  add_inplace_kernel<<<grid_for(n / 8, 256, 16), 256, 0, STREAM>>>(static_cast<__nv_bfloat16*>(a),
                                                                   static_cast<const __nv_bfloat16*>(b), n / 8);
  return launch_check("add_inplace_kernel");
}
extern "C" int eb200_copy_channels(const void* src, void* dst, long long P, int C, int scs, int scoff, int dcs,
                                   int dcoff, int accumulate, void* stream) {
  EB_REQUIRE(src && dst && C % 8 == 0, "eb200_copy_channels: bad argument");
  copy_channels_kernel<<<grid_for(P * (C / 8), 256, 16), 256, 0, STREAM>>>(
      static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), P, C, scs, scoff, dcs, dcoff,
      accumulate);
  return launch_check("copy_channels_kernel");
}
