// Inference post-processing on the GPU: what the reference runs right behind the network
// in eval mode — MT/model/postprocessing/{semantic,instance,panoptic,scene}.py and MT/utils/panoptic_merge.py
// (MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/).  The reference does this with ~30 full-tensor
// torch ops, a Python loop over the batch with .item() calls per instance (instance.py:212-266) and a CPU round trip
// for the panoptic merge (panoptic.py:140-147).  Here: a handful of HBM-bound passes, every per-instance quantity
// in small device tables, no host synchronisation.
//
// All tensors are the reference's own: fp32 NCHW network outputs, int64 index maps, uint8 instance ids.
// This translation unit is compiled WITHOUT --use_fast_math (build.py): expf / sqrtf / division are IEEE-accurate,
// because arg-max and arg-min decisions are taken on their results.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/emsanet_b200.h"
#include "common.h"

#define STREAM static_cast<cudaStream_t>(stream)
#ifndef EB200_PP_LD_DEFAULT
#define EB200_PP_LD_DEFAULT 8
#endif
#ifndef EB200_PP_V_DEFAULT
#define EB200_PP_V_DEFAULT 2
#endif

namespace {

constexpr int kMaxInst = EB200_PP_MAX_INSTANCES;   // 256 table rows: id 0 = no instance, ids 1..255 (uint8)
constexpr int kAcc = EB200_PP_ACC_FIELDS;          // per-instance accumulators, see eb200_pp_panoptic_merge

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sumf(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// aten upsample_bilinear2d (align_corners=False) source index and weights: UpSampleKernel.cpp
// compute_source_index_and_lambda / area_pixel_compute_source_index
__device__ __forceinline__ void linear_src(int dst, int in, int out, float scale, int* i0, int* i1, float* l0,
                                           float* l1) {
  if (in == out) {
    *i0 = *i1 = dst;
    *l0 = 1.f;
    *l1 = 0.f;
    return;
  }
  float s = __fsub_rn(__fmul_rn(scale, static_cast<float>(dst) + 0.5f), 0.5f);
  if (s < 0.f) s = 0.f;
  int i = static_cast<int>(floorf(s));
  if (i > in - 1) i = in - 1;
  const float l = fminf(fmaxf(s - static_cast<float>(i), 0.f), 1.f);
  *i0 = i;
  *i1 = i + (i < in - 1 ? 1 : 0);
  *l1 = l;
  *l0 = 1.f - l;
}

// ---------------------------------------------------------------------------------------------------------------
// softmax over the class axis + max / first arg-max of the softmax values, optionally on the crop of the valid
// region resampled bilinearly to (Ho, Wo): semantic.py:55-74, scene.py:41-42.  One thread per output pixel; the
// class column of a pixel is staged in shared memory ([C][128], conflict-free) so the logits are read from HBM once.
// HBM traffic = logits read once + every requested output written once.
template <bool RESAMPLE, int LD>
__global__ void __launch_bounds__(128) pp_softmax_argmax_kernel(
    const float* __restrict__ logits, int C, int H, int W, int y0, int x0, int Hc, int Wc, int Ho, int Wo, float sh,
    float sw, float* __restrict__ out_logits, float* __restrict__ scores, float* __restrict__ score,
    long long* __restrict__ idx, const unsigned char* __restrict__ cls_flags, unsigned char* __restrict__ flag_out,
    long long total) {
  extern __shared__ float col[];   // [C][128]
  const int tid = threadIdx.x;
  const long long p = static_cast<long long>(blockIdx.x) * 128 + tid;
  if (p >= total) return;          // no block-wide barrier below
  const int wo = static_cast<int>(p % Wo);
  const long long t = p / Wo;
  const int ho = static_cast<int>(t % Ho);
  const long long n = t / Ho;
  const long long HW = static_cast<long long>(H) * W;
  const long long HoWo = static_cast<long long>(Ho) * Wo;
  const float* src = logits + n * C * HW;
  const long long opix = static_cast<long long>(ho) * Wo + wo;

  // LD channels are loaded before any of them is used: LD (x4 when resampling) independent 4-byte loads in flight
  // per thread — with one load at a time the kernel sat at 41 % of the HBM roofline (Little's law, ~18 KB/SM).
  float m = -CUDART_INF_F;
  if (RESAMPLE) {
    int h0, h1, w0, w1;
    float lh0, lh1, lw0, lw1;
    linear_src(ho, Hc, Ho, sh, &h0, &h1, &lh0, &lh1);
    linear_src(wo, Wc, Wo, sw, &w0, &w1, &lw0, &lw1);
    const long long o00 = static_cast<long long>(y0 + h0) * W + x0 + w0, o01 = static_cast<long long>(y0 + h0) * W + x0 + w1;
    const long long o10 = static_cast<long long>(y0 + h1) * W + x0 + w0, o11 = static_cast<long long>(y0 + h1) * W + x0 + w1;
    constexpr int LR = LD > 1 ? LD / 4 : 1;
    for (int c0 = 0; c0 < C; c0 += LR) {
      float a[LR], b[LR], d[LR], e[LR];
#pragma unroll
      for (int j = 0; j < LR; ++j) {
        if (c0 + j < C) {
          const float* s = src + (c0 + j) * HW;
          a[j] = __ldg(s + o00);
          b[j] = __ldg(s + o01);
          d[j] = __ldg(s + o10);
          e[j] = __ldg(s + o11);
        }
      }
#pragma unroll
      for (int j = 0; j < LR; ++j) {
        if (c0 + j < C) {
          const float v = lh0 * (lw0 * a[j] + lw1 * b[j]) + lh1 * (lw0 * d[j] + lw1 * e[j]);
          col[(c0 + j) * 128 + tid] = v;
          m = fmaxf(m, v);
          if (out_logits) out_logits[(n * C + c0 + j) * HoWo + opix] = v;
        }
      }
    }
  } else {
    const long long o = static_cast<long long>(y0 + ho) * W + x0 + wo;
    for (int c0 = 0; c0 < C; c0 += LD) {
      float v[LD];
#pragma unroll
      for (int j = 0; j < LD; ++j)
        if (c0 + j < C) v[j] = __ldg(src + (c0 + j) * HW + o);
#pragma unroll
      for (int j = 0; j < LD; ++j) {
        if (c0 + j < C) {
          col[(c0 + j) * 128 + tid] = v[j];
          m = fmaxf(m, v[j]);
          if (out_logits) out_logits[(n * C + c0 + j) * HoWo + opix] = v[j];
        }
      }
    }
  }
  float sum = 0.f;
  for (int c = 0; c < C; ++c) {
    const float e = expf(col[c * 128 + tid] - m);
    col[c * 128 + tid] = e;
    sum += e;
  }
  float best = -1.f;
  int bi = 0;
  for (int c = 0; c < C; ++c) {
    const float q = col[c * 128 + tid] / sum;
    if (scores) scores[(n * C + c) * HoWo + opix] = q;
    if (q > best) {   // strict: the first maximum wins (torch.max on CPU)
      best = q;
      bi = c;
    }
  }
  if (score) score[p] = best;
  if (idx) idx[p] = bi;
  if (flag_out) flag_out[p] = cls_flags[bi] & 1;   // foreground = thing class (panoptic.py:123-128)
}

// ---------------------------------------------------------------------------------------------------------------
// Centre heat map: threshold + k x k non-maximum suppression with the reference's tie rule (instance.py:79-128).
// A pixel survives iff it is >= pad away from every border, above the threshold, equal to the maximum of its window
// and no pixel EARLIER in row-major order inside the window has the same value.  32x32 tile + halo in shared memory;
// survivors are appended (unordered) to the image's candidate list.
template <bool SKIP_EMPTY>
__global__ void __launch_bounds__(1024) pp_nms_kernel(const float* __restrict__ heat, int H, int W, int k, float thr,
                                                      float* __restrict__ cand_val, int* __restrict__ cand_idx,
                                                      int* __restrict__ cand_count, int cap) {
  extern __shared__ float tile[];
  const int pad = (k - 1) / 2, TS = 32 + 2 * pad;
  const int n = blockIdx.z;
  const float* hp = heat + static_cast<long long>(n) * H * W;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (SKIP_EMPTY) {   // most tiles of a real heat map hold nothing above the threshold: no halo load, no scan
    const int gx = bx + threadIdx.x, gy = by + threadIdx.y;
    const bool hot = gx < W && gy < H && hp[static_cast<long long>(gy) * W + gx] > thr;
    if (!__syncthreads_or(hot)) return;
  }
  for (int i = tid; i < TS * TS; i += 1024) {
    const int ty = i / TS, tx = i - ty * TS;
    const int gy = by + ty - pad, gx = bx + tx - pad;
    float v = -1.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const float u = hp[static_cast<long long>(gy) * W + gx];
      v = u > thr ? u : -1.f;   // F.threshold(x, thr, -1)
    }
    tile[i] = v;
  }
  __syncthreads();
  const int x = bx + threadIdx.x, y = by + threadIdx.y;
  if (x >= W - pad || y >= H - pad || x < pad || y < pad) return;
  const float v = tile[(threadIdx.y + pad) * TS + threadIdx.x + pad];
  if (v == -1.f) return;
  for (int dy = 0; dy < k; ++dy) {
    const float* row = tile + (threadIdx.y + dy) * TS + threadIdx.x;
    for (int dx = 0; dx < k; ++dx) {
      const float u = row[dx];
      const bool earlier = dy < pad || (dy == pad && dx < pad);
      if (u > v || (earlier && u == v)) return;
    }
  }
  const int slot = atomicAdd(&cand_count[n], 1);
  if (slot < cap) {
    cand_val[static_cast<long long>(n) * cap + slot] = v;
    cand_idx[static_cast<long long>(n) * cap + slot] = y * W + x;
  }
}

// One block per image: the top_k-th largest surviving value (bitwise search over the float bit patterns — the
// values are positive, so unsigned order = float order), clamp to >= 0, keep everything >= it (ties included,
// instance.py:131-152), order the kept centres row-major (nonzero(), :158-159) = instance ids 1..K.
__global__ void __launch_bounds__(1024) pp_select_kernel(const float* __restrict__ cand_val,
                                                         const int* __restrict__ cand_idx,
                                                         const int* __restrict__ cand_count, int cap, int top_k, int W,
                                                         long long HW, const unsigned char* __restrict__ fg,
                                                         int* __restrict__ centers, float* __restrict__ cscore,
                                                         int* __restrict__ ccount, int* __restrict__ status) {
  __shared__ int s_cnt, s_keep;
  __shared__ float k_val[1024];
  __shared__ int k_idx[1024];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int found = cand_count[n];
  const int M = found < cap ? found : cap;
  int st = found > cap ? 1 : 0;
  const float* v = cand_val + static_cast<long long>(n) * cap;
  const int* ix = cand_idx + static_cast<long long>(n) * cap;
  float lowest = 0.f;   // fewer than top_k survivors: the k-th value is -1, clamped to 0 (instance.py:146)
  if (M >= top_k) {
    unsigned prefix = 0;
    for (int bit = 30; bit >= 0; --bit) {
      const unsigned trial = prefix | (1u << bit);
      if (tid == 0) s_cnt = 0;
      __syncthreads();
      int c = 0;
      for (int i = tid; i < M; i += 1024) c += (__float_as_uint(v[i]) >= trial) ? 1 : 0;
      c = warp_sum(c);
      if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
      __syncthreads();
      if (s_cnt >= top_k) prefix = trial;
      __syncthreads();
    }
    lowest = __uint_as_float(prefix);
  }
  if (tid == 0) s_keep = 0;
  __syncthreads();
  for (int i = tid; i < M; i += 1024) {
    const float x = v[i];
    const int id = ix[i];
    if (x >= lowest && (fg == nullptr || fg[static_cast<long long>(n) * HW + id])) {
      const int s = atomicAdd(&s_keep, 1);
      if (s < 1024) {
        k_val[s] = x;
        k_idx[s] = id;
      }
    }
  }
  __syncthreads();
  const int K = s_keep;
  if (K > kMaxInst - 1) st |= 2;          // more centres than uint8 ids: the reference would wrap around
  const int Kc = K < 1024 ? K : 1024;
  for (int i = tid; i < Kc; i += 1024) {
    const int mine = k_idx[i];
    int r = 0;
    for (int j = 0; j < Kc; ++j) r += (k_idx[j] < mine) ? 1 : 0;
    if (r < kMaxInst - 1) {
      centers[(static_cast<long long>(n) * kMaxInst + r) * 2 + 0] = mine / W;
      centers[(static_cast<long long>(n) * kMaxInst + r) * 2 + 1] = mine % W;
      cscore[static_cast<long long>(n) * kMaxInst + r] = k_val[i];
    }
  }
  if (tid == 0) {
    ccount[n] = Kc < kMaxInst - 1 ? Kc : kMaxInst - 1;
    status[n] = st;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pixel -> instance id (instance.py:176-253): id = 1 + first arg-min over the centres of
// || centre - (pixel + offset) ||_2, fp32, evaluated exactly as written (no FMA contraction); areas (bincount) and,
// for the panoptic merge, the votes[id][semantic class + 1] histogram (panoptic_merge.py:196-199).
template <bool LAZY_SQRT>
__global__ void __launch_bounds__(256) pp_assign_kernel(const float* __restrict__ offset,
                                                        const unsigned char* __restrict__ fg,
                                                        const int* __restrict__ centers, const int* __restrict__ ccount,
                                                        int H, int W, float scale_y, float scale_x, float dist_thr,
                                                        const long long* __restrict__ sem_idx, int Cp1,
                                                        unsigned char* __restrict__ seg, int* __restrict__ areas,
                                                        int* __restrict__ votes) {
  __shared__ float cy[kMaxInst], cx[kMaxInst];
  __shared__ int s_area[kMaxInst];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int K = ccount[n];
  const long long HW = static_cast<long long>(H) * W;
  s_area[tid] = 0;
  if (tid < K) {
    cy[tid] = static_cast<float>(centers[(static_cast<long long>(n) * kMaxInst + tid) * 2 + 0]);
    cx[tid] = static_cast<float>(centers[(static_cast<long long>(n) * kMaxInst + tid) * 2 + 1]);
  }
  __syncthreads();
  const long long p = static_cast<long long>(blockIdx.x) * 256 + tid;
  const bool valid = p < HW;
  int id = 0;
  if (valid && K > 0 && fg[n * HW + p]) {
    const int y = static_cast<int>(p / W), x = static_cast<int>(p - static_cast<long long>(y) * W);
    const float ly = __fadd_rn(static_cast<float>(y), __fmul_rn(offset[(n * 2 + 0) * HW + p], scale_y));
    const float lx = __fadd_rn(static_cast<float>(x), __fmul_rn(offset[(n * 2 + 1) * HW + p], scale_x));
    float best = CUDART_INF_F;
    int bi = 0;
    if (LAZY_SQRT) {
      // sqrt is monotone, so min_j sqrt(d2_j) = sqrt(min_j d2_j) and the reference's answer — the FIRST j whose
      // ROUNDED distance equals that minimum — can only be a j whose d2_j lies within float rounding of the smallest
      // d2 (two d2 values merge under sqrtf only if they differ by < 2.4e-7 relative): one sqrt per pixel plus one
      // per near-tie instead of one per centre.
      float m2 = CUDART_INF_F;
      for (int j = 0; j < K; ++j) {
        const float dy = __fsub_rn(cy[j], ly), dx = __fsub_rn(cx[j], lx);
        m2 = fminf(m2, __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
      }
      best = sqrtf(m2);
      const float lim = __fmul_rn(m2, 1.000001f);
      for (int j = 0; j < K; ++j) {
        const float dy = __fsub_rn(cy[j], ly), dx = __fsub_rn(cx[j], lx);
        const float d2 = __fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx));
        if (d2 <= lim && sqrtf(d2) == best) {
          bi = j;
          break;
        }
      }
    } else {
      for (int j = 0; j < K; ++j) {
        const float dy = __fsub_rn(cy[j], ly), dx = __fsub_rn(cx[j], lx);
        const float d = sqrtf(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
        if (d < best) {   // strict: the first minimum wins (torch.min on CPU)
          best = d;
          bi = j;
        }
      }
    }
    id = bi + 1;
    if (dist_thr >= 0.f && best > dist_thr) id = 0;   // instance.py:236-238
    atomicAdd(&s_area[id], 1);
  }
  if (valid) seg[n * HW + p] = static_cast<unsigned char>(id);
  // votes: neighbouring pixels mostly share (id, class) — one atomic per distinct key per warp
  const bool vote = votes != nullptr && id > 0;
  const unsigned voters = __ballot_sync(0xffffffffu, vote);
  if (vote) {
    const int key = id * Cp1 + static_cast<int>(sem_idx[n * HW + p]) + 1;
    const unsigned same = __match_any_sync(voters, key);
    if ((tid & 31) == __ffs(same) - 1)
      atomicAdd(&votes[static_cast<long long>(n) * kMaxInst * Cp1 + key], __popc(same));
  }
  __syncthreads();
  if (s_area[tid]) atomicAdd(&areas[static_cast<long long>(n) * kMaxInst + tid], s_area[tid]);
}

// ---------------------------------------------------------------------------------------------------------------
// Panoptic merge (panoptic_merge.py:168-225).  Table pass, one block per image: class of an instance = the
// smallest most frequent semantic label of its pixels (torch.mode), new id = running count of that class in
// ascending instance-id order, panoptic id = class * 2^16 + new id.
__global__ void __launch_bounds__(kMaxInst) pp_merge_table_kernel(const int* __restrict__ votes,
                                                                   const int* __restrict__ ccount, int Cp1,
                                                                   int* __restrict__ inst_pan) {
  extern __shared__ int tracker[];   // [Cp1]
  __shared__ int cls_of[kMaxInst];
  const int n = blockIdx.x, id = threadIdx.x;
  const int K = ccount[n];
  int cls = 0;
  if (id >= 1 && id <= K) {
    const int* v = votes + (static_cast<long long>(n) * kMaxInst + id) * Cp1;
    int best = 0;
    for (int c = 0; c < Cp1; ++c) {
      const int q = v[c];
      if (q > best) {
        best = q;
        cls = c;
      }
    }                                 // no pixels: best stays 0 -> cls 0 -> skipped like the reference (:192-193)
  }
  cls_of[id] = cls;
  for (int c = id; c < Cp1; c += kMaxInst) tracker[c] = 0;
  if (id == 0 || id > K) inst_pan[static_cast<long long>(n) * kMaxInst + id] = 0;
  __syncthreads();
  if (id == 0) {
    for (int i = 1; i <= K; ++i) {
      const int c = cls_of[i];
      inst_pan[static_cast<long long>(n) * kMaxInst + i] = c ? c * 65536 + (++tracker[c]) : 0;
    }
  }
}

// Pixel pass: panoptic id, its semantic class, the semantic score of that class (panoptic.py:150-190) and the
// per-instance accumulators {sum of semantic score, pixels, sum cos, sum sin, oriented pixels} (:204-216, :289-303
// with instance.py:275-323).  cls_flags[c]: bit 0 = thing class, bit 1 = class has orientation.
__global__ void __launch_bounds__(256) pp_panoptic_kernel(
    const unsigned char* __restrict__ seg, const long long* __restrict__ sem_idx,
    const unsigned char* __restrict__ cls_flags, const int* __restrict__ inst_pan, const float* __restrict__ scores,
    const float* __restrict__ orient, int C, long long HW, long long* __restrict__ pan,
    long long* __restrict__ pan_sem, float* __restrict__ sem_score, double* __restrict__ inst_acc) {
  __shared__ float acc[kMaxInst * kAcc];
  const int n = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < kMaxInst * kAcc; i += 256) acc[i] = 0.f;
  __syncthreads();
  const long long p = static_cast<long long>(blockIdx.x) * 256 + tid;
  const bool valid = p < HW;
  int id = 0, pid = 0;
  float sc = 0.f, oc = 0.f, os = 0.f, on = 0.f;
  if (valid) {
    id = seg[n * HW + p];
    const int s = static_cast<int>(sem_idx[n * HW + p]);
    pid = id > 0 ? inst_pan[static_cast<long long>(n) * kMaxInst + id] : ((cls_flags[s] & 1) ? 0 : (s + 1) << 16);
    const int ps = pid >> 16;
    pan[n * HW + p] = pid;
    pan_sem[n * HW + p] = ps;
    if (scores) {
      sc = ps > 0 ? scores[(n * C + ps - 1) * HW + p] : 0.f;
      sem_score[n * HW + p] = sc;
    }
    if (orient && id > 0 && ps > 0 && (cls_flags[ps - 1] & 2)) {
      oc = orient[(n * 2 + 0) * HW + p];
      os = orient[(n * 2 + 1) * HW + p];
      on = 1.f;
    }
  }
  // a warp that lies inside one instance (the common case) reduces with shuffles and adds once
  const int id0 = __shfl_sync(0xffffffffu, id, 0);
  const bool uniform = __all_sync(0xffffffffu, id == id0);
  const float one = (id > 0 && pid > 0) ? 1.f : 0.f;
  if (uniform) {
    if (id0 > 0) {
      const float a0 = warp_sumf(one * sc), a1 = warp_sumf(one), a2 = warp_sumf(oc), a3 = warp_sumf(os),
                  a4 = warp_sumf(on);
      if ((tid & 31) == 0) {
        float* a = acc + id0 * kAcc;
        atomicAdd(a + 0, a0);
        atomicAdd(a + 1, a1);
        atomicAdd(a + 2, a2);
        atomicAdd(a + 3, a3);
        atomicAdd(a + 4, a4);
      }
    }
  } else if (one > 0.f) {
    float* a = acc + id * kAcc;
    atomicAdd(a + 0, sc);
    atomicAdd(a + 1, 1.f);
    if (on > 0.f) {
      atomicAdd(a + 2, oc);
      atomicAdd(a + 3, os);
      atomicAdd(a + 4, 1.f);
    }
  }
  __syncthreads();
  for (int i = tid; i < kMaxInst * kAcc; i += 256)
    if (acc[i] != 0.f) atomicAdd(&inst_acc[static_cast<long long>(n) * kMaxInst * kAcc + i], static_cast<double>(acc[i]));
}

// The same pass with FOUR consecutive pixels per thread (HW % 4 == 0): 16/32-byte vector loads and stores, a quarter
// of the blocks (the one-pixel version spent its time on per-block set-up: zeroing / scanning the 5 KB accumulator
// and two barriers per 256 pixels — 1.0 TB/s in ncu), one shuffle reduction per 128 pixels.
__global__ void __launch_bounds__(256) pp_panoptic4_kernel(
    const unsigned char* __restrict__ seg, const long long* __restrict__ sem_idx,
    const unsigned char* __restrict__ cls_flags, const int* __restrict__ inst_pan, const float* __restrict__ scores,
    const float* __restrict__ orient, int C, long long HW, long long* __restrict__ pan,
    long long* __restrict__ pan_sem, float* __restrict__ sem_score, double* __restrict__ inst_acc) {
  __shared__ float acc[kMaxInst * kAcc];
  const int n = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < kMaxInst * kAcc; i += 256) acc[i] = 0.f;
  __syncthreads();
  const long long p0 = (static_cast<long long>(blockIdx.x) * 256 + tid) * 4;
  const bool valid = p0 < HW;   // HW % 4 == 0: all four pixels or none
  int id[4] = {0, 0, 0, 0}, pid[4] = {0, 0, 0, 0};
  float sc[4] = {0.f, 0.f, 0.f, 0.f}, oc[4] = {0.f, 0.f, 0.f, 0.f}, os[4] = {0.f, 0.f, 0.f, 0.f},
        on[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const long long g = n * HW + p0;
    const uchar4 s4 = *reinterpret_cast<const uchar4*>(seg + g);
    const longlong2 ca = *reinterpret_cast<const longlong2*>(sem_idx + g);
    const longlong2 cb = *reinterpret_cast<const longlong2*>(sem_idx + g + 2);
    id[0] = s4.x, id[1] = s4.y, id[2] = s4.z, id[3] = s4.w;
    const int cls[4] = {static_cast<int>(ca.x), static_cast<int>(ca.y), static_cast<int>(cb.x), static_cast<int>(cb.y)};
    int ps[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pid[q] = id[q] > 0 ? inst_pan[static_cast<long long>(n) * kMaxInst + id[q]]
                         : ((cls_flags[cls[q]] & 1) ? 0 : (cls[q] + 1) << 16);
      ps[q] = pid[q] >> 16;
    }
    *reinterpret_cast<longlong2*>(pan + g) = make_longlong2(pid[0], pid[1]);
    *reinterpret_cast<longlong2*>(pan + g + 2) = make_longlong2(pid[2], pid[3]);
    *reinterpret_cast<longlong2*>(pan_sem + g) = make_longlong2(ps[0], ps[1]);
    *reinterpret_cast<longlong2*>(pan_sem + g + 2) = make_longlong2(ps[2], ps[3]);
    if (scores) {
#pragma unroll
      for (int q = 0; q < 4; ++q) sc[q] = ps[q] > 0 ? scores[(static_cast<long long>(n) * C + ps[q] - 1) * HW + p0 + q] : 0.f;
      *reinterpret_cast<float4*>(sem_score + g) = make_float4(sc[0], sc[1], sc[2], sc[3]);
    }
    if (orient) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (id[q] > 0 && ps[q] > 0 && (cls_flags[ps[q] - 1] & 2)) {
          oc[q] = orient[(static_cast<long long>(n) * 2 + 0) * HW + p0 + q];
          os[q] = orient[(static_cast<long long>(n) * 2 + 1) * HW + p0 + q];
          on[q] = 1.f;
        }
      }
    }
  }
  // a warp whose 128 pixels lie inside one instance reduces with shuffles and adds once
  const int id0 = __shfl_sync(0xffffffffu, id[0], 0);
  const bool mine = id[0] == id0 && id[1] == id0 && id[2] == id0 && id[3] == id0;
  const bool uniform = __all_sync(0xffffffffu, mine);
  if (uniform) {
    if (id0 > 0) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float one = pid[q] > 0 ? 1.f : 0.f;
        t0 += one * sc[q];
        t1 += one;
        t2 += oc[q];
        t3 += os[q];
        t4 += on[q];
      }
      t0 = warp_sumf(t0), t1 = warp_sumf(t1), t2 = warp_sumf(t2), t3 = warp_sumf(t3), t4 = warp_sumf(t4);
      if ((tid & 31) == 0) {
        float* a = acc + id0 * kAcc;
        atomicAdd(a + 0, t0);
        atomicAdd(a + 1, t1);
        atomicAdd(a + 2, t2);
        atomicAdd(a + 3, t3);
        atomicAdd(a + 4, t4);
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (id[q] > 0 && pid[q] > 0) {
        float* a = acc + id[q] * kAcc;
        atomicAdd(a + 0, sc[q]);
        atomicAdd(a + 1, 1.f);
        if (on[q] > 0.f) {
          atomicAdd(a + 2, oc[q]);
          atomicAdd(a + 3, os[q]);
          atomicAdd(a + 4, 1.f);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < kMaxInst * kAcc; i += 256)
    if (acc[i] != 0.f) atomicAdd(&inst_acc[static_cast<long long>(n) * kMaxInst * kAcc + i], static_cast<double>(acc[i]));
}

// Score maps (panoptic.py:192-236): instance score = the centre's heat value, panoptic score = mean semantic score
// of the instance * instance score on instance pixels, the semantic score elsewhere.
__global__ void __launch_bounds__(256) pp_score_maps_kernel(const unsigned char* __restrict__ seg,
                                                            const int* __restrict__ inst_pan,
                                                            const float* __restrict__ cscore,
                                                            const double* __restrict__ inst_acc,
                                                            const float* __restrict__ sem_score, long long HW,
                                                            float* __restrict__ ins_score,
                                                            float* __restrict__ pan_score) {
  const int n = blockIdx.y;
  const long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= HW) return;
  const int id = seg[n * HW + p];
  float is = 0.f, ps = sem_score[n * HW + p];
  if (id > 0 && inst_pan[static_cast<long long>(n) * kMaxInst + id] > 0) {
    const double* a = inst_acc + (static_cast<long long>(n) * kMaxInst + id) * kAcc;
    is = cscore[static_cast<long long>(n) * kMaxInst + id - 1];
    ps = __fmul_rn(static_cast<float>(a[0] / a[1]), is);
  }
  ins_score[n * HW + p] = is;
  pan_score[n * HW + p] = ps;
}

// aten upsample_nearest2d on the crop of the valid region (dense_base.py:15-58, mode='nearest')
template <typename T>
__global__ void __launch_bounds__(256) pp_nearest_kernel(const T* __restrict__ in, int H, int W, int y0, int x0, int Hc,
                                                         int Wc, T* __restrict__ out, int Ho, int Wo, float sh,
                                                         float sw, long long total) {
  const long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (p >= total) return;
  const int wo = static_cast<int>(p % Wo);
  const long long t = p / Wo;
  const int ho = static_cast<int>(t % Ho);
  const long long n = t / Ho;
  int hs = Ho == Hc ? ho : (Ho == 2 * Hc ? ho >> 1 : static_cast<int>(floorf(__fmul_rn(static_cast<float>(ho), sh))));
  int ws = Wo == Wc ? wo : (Wo == 2 * Wc ? wo >> 1 : static_cast<int>(floorf(__fmul_rn(static_cast<float>(wo), sw))));
  if (hs > Hc - 1) hs = Hc - 1;
  if (ws > Wc - 1) ws = Wc - 1;
  out[p] = in[(n * H + y0 + hs) * W + x0 + ws];
}

// EB200_PP_V: 1 = first kernels (one pixel per thread, sqrt per centre), 2 = the restructured ones
inline int pp_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EB200_PP_V");
    v = e ? atoi(e) : EB200_PP_V_DEFAULT;
  }
  return v;
}

// Per-instance orientation on an ARBITRARY instance map (InstancePostprocessing._get_instance_orientation,
// instance.py:275-323, called with ground-truth maps in dataset evaluation, :431-451): acc[n][id] += (cos, sin, 1) over
// the pixels with seg == id > 0 inside the foreground mask.  seg elements of 1 / 2 / 4 / 8 bytes.
__device__ __forceinline__ long long load_id(const void* seg, int bytes, long long i) {
  switch (bytes) {
    case 1: return static_cast<const unsigned char*>(seg)[i];
    case 2: return static_cast<const unsigned short*>(seg)[i];
    case 4: return static_cast<const int*>(seg)[i];
    default: return static_cast<const long long*>(seg)[i];
  }
}

__global__ void __launch_bounds__(256) pp_orientation_acc_kernel(const float* __restrict__ orient,
                                                                 const void* __restrict__ seg, int seg_bytes,
                                                                 const unsigned char* __restrict__ fg, long long HW,
                                                                 int max_id, double* __restrict__ acc) {
  const int n = blockIdx.y;
  const long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  long long id = 0;
  float c = 0.f, s = 0.f;
  if (p < HW) {
    id = load_id(seg, seg_bytes, n * HW + p);
    if (id < 0 || id > max_id || (fg != nullptr && !fg[n * HW + p])) id = 0;
    if (id > 0) {
      c = orient[(static_cast<long long>(n) * 2 + 0) * HW + p];
      s = orient[(static_cast<long long>(n) * 2 + 1) * HW + p];
    }
  }
  const long long id0 = __shfl_sync(0xffffffffu, id, 0);
  const bool uniform = __all_sync(0xffffffffu, id == id0);
  double* a = acc + (static_cast<long long>(n) * (max_id + 1) + id) * 3;
  if (uniform) {
    if (id0 > 0) {
      double dc = c, ds = s;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        dc += __shfl_xor_sync(0xffffffffu, dc, o);
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(a + 0, dc);
        atomicAdd(a + 1, ds);
        atomicAdd(a + 2, 32.0);
      }
    }
  } else if (id > 0) {
    atomicAdd(a + 0, static_cast<double>(c));
    atomicAdd(a + 1, static_cast<double>(s));
    atomicAdd(a + 2, 1.0);
  }
}

inline int blocks_for(long long items, int per_block) {
  return static_cast<int>((items + per_block - 1) / per_block);
}

}  // namespace

// ===============================================================================================================
extern "C" int eb200_pp_softmax_argmax(const float* logits, int N, int C, int H, int W, int y0, int x0, int Hc, int Wc,
                                       int Ho, int Wo, float* out_logits, float* scores, float* score, long long* idx,
                                       const unsigned char* cls_flags, unsigned char* flag_out, void* stream) {
  EB_REQUIRE(logits && N > 0 && C > 0 && H > 0 && W > 0, "eb200_pp_softmax_argmax: bad argument");
  EB_REQUIRE(y0 >= 0 && x0 >= 0 && Hc > 0 && Wc > 0 && y0 + Hc <= H && x0 + Wc <= W && Ho > 0 && Wo > 0,
             "eb200_pp_softmax_argmax: crop window (%d,%d,%d,%d) outside %dx%d", y0, x0, Hc, Wc, H, W);
  EB_REQUIRE(!flag_out || cls_flags, "eb200_pp_softmax_argmax: flag_out needs cls_flags");
  const size_t smem = static_cast<size_t>(C) * 128 * sizeof(float);
  EB_REQUIRE(smem <= 200 * 1024, "eb200_pp_softmax_argmax: %d classes exceed the shared-memory column buffer", C);
  const long long total = static_cast<long long>(N) * Ho * Wo;
  const bool resample = !(Ho == Hc && Wo == Wc);
  const float sh = static_cast<float>(Hc) / static_cast<float>(Ho), sw = static_cast<float>(Wc) / static_cast<float>(Wo);
  static int ld = -1;   // EB200_PP_LD: channels loaded ahead per thread (1 = one load at a time, 8 = default)
  if (ld < 0) {
    const char* e = getenv("EB200_PP_LD");
    ld = e ? atoi(e) : EB200_PP_LD_DEFAULT;
  }
  auto fn = ld > 1 ? (resample ? pp_softmax_argmax_kernel<true, 8> : pp_softmax_argmax_kernel<false, 8>)
                   : (resample ? pp_softmax_argmax_kernel<true, 1> : pp_softmax_argmax_kernel<false, 1>);
  if (smem > 48 * 1024)
    EB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  fn<<<blocks_for(total, 128), 128, smem, STREAM>>>(logits, C, H, W, y0, x0, Hc, Wc, Ho, Wo, sh, sw, out_logits, scores,
                                                    score, idx, cls_flags, flag_out, total);
  return eb::launch_check("pp_softmax_argmax_kernel");
}

extern "C" long long eb200_pp_centers_ws_bytes(int N, int H, int W, int nms_k) {
  const int pad = (nms_k - 1) / 2;
  // two survivors are never inside each other's window (the tie rule removes the later one)
  const long long cap = static_cast<long long>((H + pad) / (pad + 1)) * ((W + pad) / (pad + 1));
  return static_cast<long long>(N) * (cap * 8 + 4) + 256;
}

extern "C" int eb200_pp_instance_centers(const float* heat, int N, int H, int W, float threshold, int nms_k, int top_k,
                                         const unsigned char* fg, void* ws, long long ws_bytes, int* centers,
                                         float* center_scores, int* counts, int* status, void* stream) {
  EB_REQUIRE(heat && ws && centers && center_scores && counts && status && N > 0, "eb200_pp_instance_centers: bad argument");
  EB_REQUIRE(nms_k >= 3 && nms_k % 2 == 1 && nms_k <= 33, "eb200_pp_instance_centers: nms kernel %d not odd in [3,33]", nms_k);
  EB_REQUIRE(threshold >= 0.f, "eb200_pp_instance_centers: negative heat-map threshold");
  EB_REQUIRE(top_k >= 1 && top_k <= 254, "eb200_pp_instance_centers: top_k %d outside [1,254] (instance.py:38)", top_k);
  EB_REQUIRE(ws_bytes >= eb200_pp_centers_ws_bytes(N, H, W, nms_k), "eb200_pp_instance_centers: workspace too small");
  const int pad = (nms_k - 1) / 2;
  const long long cap = static_cast<long long>((H + pad) / (pad + 1)) * ((W + pad) / (pad + 1));
  EB_REQUIRE(cap < (1ll << 30), "eb200_pp_instance_centers: image too large");
  int* cand_count = static_cast<int*>(ws);
  float* cand_val = reinterpret_cast<float*>(static_cast<char*>(ws) + ((static_cast<long long>(N) * 4 + 255) / 256) * 256);
  int* cand_idx = reinterpret_cast<int*>(cand_val + static_cast<long long>(N) * cap);
  EB_CUDA(cudaMemsetAsync(cand_count, 0, static_cast<size_t>(N) * 4, STREAM));
  const int TS = 32 + 2 * pad;
  dim3 grid(eb::ceil_div(W, 32), eb::ceil_div(H, 32), N);
  auto nms = pp_version() >= 2 ? pp_nms_kernel<true> : pp_nms_kernel<false>;
  nms<<<grid, dim3(32, 32), static_cast<size_t>(TS) * TS * 4, STREAM>>>(heat, H, W, nms_k, threshold, cand_val, cand_idx,
                                                                        cand_count, static_cast<int>(cap));
  if (int rc = eb::launch_check("pp_nms_kernel")) return rc;
  pp_select_kernel<<<N, 1024, 0, STREAM>>>(cand_val, cand_idx, cand_count, static_cast<int>(cap), top_k, W,
                                           static_cast<long long>(H) * W, fg, centers, center_scores, counts, status);
  return eb::launch_check("pp_select_kernel");
}

extern "C" int eb200_pp_instance_assign(const float* offset, const unsigned char* fg, const int* centers,
                                        const int* counts, int N, int H, int W, float scale_y, float scale_x,
                                        float dist_thr, const long long* sem_idx, int n_classes, unsigned char* seg,
                                        int* areas, int* votes, void* stream) {
  EB_REQUIRE(offset && fg && centers && counts && seg && areas && N > 0, "eb200_pp_instance_assign: bad argument");
  EB_REQUIRE(!votes || (sem_idx && n_classes > 0), "eb200_pp_instance_assign: votes need sem_idx and n_classes");
  const int Cp1 = n_classes + 1;
  EB_CUDA(cudaMemsetAsync(areas, 0, static_cast<size_t>(N) * kMaxInst * 4, STREAM));
  if (votes) EB_CUDA(cudaMemsetAsync(votes, 0, static_cast<size_t>(N) * kMaxInst * Cp1 * 4, STREAM));
  dim3 grid(blocks_for(static_cast<long long>(H) * W, 256), N);
  auto assign = pp_version() >= 2 ? pp_assign_kernel<true> : pp_assign_kernel<false>;
  assign<<<grid, 256, 0, STREAM>>>(offset, fg, centers, counts, H, W, scale_y, scale_x, dist_thr, sem_idx, Cp1, seg, areas,
                                   votes);
  return eb::launch_check("pp_assign_kernel");
}

extern "C" int eb200_pp_panoptic_merge(const unsigned char* seg, const long long* sem_idx,
                                       const unsigned char* cls_flags, const int* counts, const int* votes,
                                       const float* scores, const float* orientation, const float* center_scores, int N,
                                       int C, int H, int W, int* inst_pan, long long* pan, long long* pan_sem,
                                       float* sem_score, float* ins_score, float* pan_score, double* inst_acc,
                                       void* stream) {
  EB_REQUIRE(seg && sem_idx && cls_flags && counts && votes && inst_pan && pan && pan_sem && inst_acc && N > 0 && C > 0,
             "eb200_pp_panoptic_merge: bad argument");
  EB_REQUIRE(!scores || (sem_score && ins_score && pan_score && center_scores),
             "eb200_pp_panoptic_merge: score maps need sem_score, ins_score, pan_score and center_scores");
  const long long HW = static_cast<long long>(H) * W;
  EB_CUDA(cudaMemsetAsync(inst_acc, 0, static_cast<size_t>(N) * kMaxInst * kAcc * sizeof(double), STREAM));
  pp_merge_table_kernel<<<N, kMaxInst, static_cast<size_t>(C + 1) * 4, STREAM>>>(votes, counts, C + 1, inst_pan);
  if (int rc = eb::launch_check("pp_merge_table_kernel")) return rc;
  dim3 grid(blocks_for(HW, 256), N);
  const bool aligned = ((reinterpret_cast<uintptr_t>(seg) & 3) | (reinterpret_cast<uintptr_t>(sem_idx) & 15) |
                        (reinterpret_cast<uintptr_t>(pan) & 15) | (reinterpret_cast<uintptr_t>(pan_sem) & 15) |
                        (reinterpret_cast<uintptr_t>(sem_score) & 15)) == 0;
  if (pp_version() >= 2 && HW % 4 == 0 && aligned) {
    dim3 grid4(blocks_for(HW, 1024), N);
    pp_panoptic4_kernel<<<grid4, 256, 0, STREAM>>>(seg, sem_idx, cls_flags, inst_pan, scores, orientation, C, HW, pan,
                                                   pan_sem, sem_score, inst_acc);
  } else {
    pp_panoptic_kernel<<<grid, 256, 0, STREAM>>>(seg, sem_idx, cls_flags, inst_pan, scores, orientation, C, HW, pan,
                                                 pan_sem, sem_score, inst_acc);
  }
  if (int rc = eb::launch_check("pp_panoptic_kernel")) return rc;
  if (scores) {
    pp_score_maps_kernel<<<grid, 256, 0, STREAM>>>(seg, inst_pan, center_scores, inst_acc, sem_score, HW, ins_score,
                                                   pan_score);
    return eb::launch_check("pp_score_maps_kernel");
  }
  return 0;
}

extern "C" int eb200_pp_nearest_resize(const void* in, void* out, int elem_bytes, int N, int H, int W, int y0, int x0,
                                       int Hc, int Wc, int Ho, int Wo, void* stream) {
  EB_REQUIRE(in && out && N > 0, "eb200_pp_nearest_resize: bad argument");
  EB_REQUIRE(y0 >= 0 && x0 >= 0 && Hc > 0 && Wc > 0 && y0 + Hc <= H && x0 + Wc <= W && Ho > 0 && Wo > 0,
             "eb200_pp_nearest_resize: crop window outside the image");
  const long long total = static_cast<long long>(N) * Ho * Wo;
  const float sh = static_cast<float>(Hc) / static_cast<float>(Ho), sw = static_cast<float>(Wc) / static_cast<float>(Wo);
  const int grid = blocks_for(total, 256);
  switch (elem_bytes) {
    case 1:
      pp_nearest_kernel<unsigned char><<<grid, 256, 0, STREAM>>>(static_cast<const unsigned char*>(in), H, W, y0, x0, Hc,
                                                                 Wc, static_cast<unsigned char*>(out), Ho, Wo, sh, sw,
                                                                 total);
      break;
    case 4:
      pp_nearest_kernel<unsigned int><<<grid, 256, 0, STREAM>>>(static_cast<const unsigned int*>(in), H, W, y0, x0, Hc, Wc,
                                                                static_cast<unsigned int*>(out), Ho, Wo, sh, sw, total);
      break;
    case 8:
      pp_nearest_kernel<unsigned long long><<<grid, 256, 0, STREAM>>>(static_cast<const unsigned long long*>(in), H, W, y0,
                                                                      x0, Hc, Wc, static_cast<unsigned long long*>(out),
                                                                      Ho, Wo, sh, sw, total);
      break;
    default:
      return eb::fail("eb200_pp_nearest_resize: element size %d not in {1,4,8}", elem_bytes);
  }
  return eb::launch_check("pp_nearest_kernel");
}

extern "C" int eb200_pp_instance_orientation(const float* orientation, const void* seg, int seg_bytes,
                                             const unsigned char* fg, int N, int H, int W, int max_id, double* acc,
                                             void* stream) {
  EB_REQUIRE(orientation && seg && acc && N > 0 && H > 0 && W > 0 && max_id >= 0,
             "eb200_pp_instance_orientation: bad argument");
  EB_REQUIRE(seg_bytes == 1 || seg_bytes == 2 || seg_bytes == 4 || seg_bytes == 8,
             "eb200_pp_instance_orientation: instance ids of %d bytes", seg_bytes);
  const long long HW = static_cast<long long>(H) * W;
  EB_CUDA(cudaMemsetAsync(acc, 0, static_cast<size_t>(N) * (max_id + 1) * 3 * sizeof(double), STREAM));
  dim3 grid(blocks_for(HW, 256), N);
  pp_orientation_acc_kernel<<<grid, 256, 0, STREAM>>>(orientation, seg, seg_bytes, fg, HW, max_id, acc);
  return eb::launch_check("pp_orientation_acc_kernel");
}
