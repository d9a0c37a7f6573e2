// Fused semantic cross-entropy: MT/loss/ce.py:13-68 with weighted_reduction=False, i.e.
// torch.nn.CrossEntropyLoss(weight, reduction='sum', ignore_index=-1, label_smoothing) on `target - 1`
// (MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/).  The reference runs log_softmax, nll_loss,
// the smoothing term and their autograd backward as separate full-tensor passes over the network's largest output
// (N x 40 x 480 x 640 fp32 = 1.57 GB at the bench batch: ~8 GB of HBM traffic per step) plus a host synchronisation
// for the element count.  Here: the forward reads the logits once (loss and count accumulated on the device), the
// backward reads them once more and writes the gradient, already scaled by the upstream gradient: 3 tensor passes.
//
// Compiled without --use_fast_math (build.py): expf / logf / division are IEEE-accurate.
#include <cuda_runtime.h>

#include <algorithm>
#include <math_constants.h>

#include "../../include/emsanet_b200.h"
#include "common.h"

#define STREAM static_cast<cudaStream_t>(stream)

namespace {

__device__ __forceinline__ int load_target(const void* target, int target_bytes, long long i) {
  // the reference shifts by -1: 0 = void -> ignore_index (ce.py:46)
  if (target_bytes == 1) return static_cast<int>(static_cast<const unsigned char*>(target)[i]) - 1;
  if (target_bytes == 4) return static_cast<const int*>(target)[i] - 1;
  return static_cast<int>(static_cast<const long long*>(target)[i]) - 1;
}

// One thread per pixel; the pixel's class column is staged in shared memory ([C][128]) with 8 loads in flight.
// BACKWARD = false: loss_i summed per CTA in fp64, one fp64 atomic per CTA; non-void pixels counted.
// BACKWARD = true : dlogits[k] = grad_out * (-a[k] + (sum_c a[c]) * softmax[k]), zeros on void pixels.
template <bool BACKWARD>
__global__ void __launch_bounds__(128) ce_kernel(const float* __restrict__ logits, const void* __restrict__ target,
                                                 int target_bytes, const float* __restrict__ weights, float eps, int C,
                                                 long long HW, long long total, const float* __restrict__ grad_out,
                                                 float* __restrict__ dlogits, double* __restrict__ loss_acc,
                                                 long long* __restrict__ count_acc) {
  extern __shared__ float col[];   // [C][128]
  __shared__ double s_loss[4];
  __shared__ int s_cnt[4];
  const int tid = threadIdx.x;
  const long long p = static_cast<long long>(blockIdx.x) * 128 + tid;
  const bool in_range = p < total;
  double my_loss = 0.0;
  int my_cnt = 0;
  if (in_range) {
    const long long n = p / HW, pix = p - n * HW;
    const float* src = logits + n * C * HW + pix;
    const int t = load_target(target, target_bytes, p);
    const bool valid = t >= 0 && t < C;
    // a label >= C is a dataset / n_classes mismatch: torch's CrossEntropyLoss device-asserts on it.  Here it contributes
    // no loss, is COUNTED like the reference counts it (ce.py:50: target_shifted >= 0) and raises the out-of-range flag
    // (bit 62 of the count word), which the host side turns into an error.
    if (!BACKWARD && t >= C) my_cnt = 1 | (1 << 20);
    if (valid || BACKWARD) {
      if (!valid) {
        for (int c = 0; c < C; ++c) dlogits[(n * C + c) * HW + pix] = 0.f;
      } else {
        float m = -CUDART_INF_F;
        for (int c0 = 0; c0 < C; c0 += 8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c0 + j < C) v[j] = __ldg(src + (c0 + j) * HW);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (c0 + j < C) {
              col[(c0 + j) * 128 + tid] = v[j];
              m = fmaxf(m, v[j]);
            }
          }
        }
        float sum = 0.f, wx = 0.f, wsum = 0.f;   // sum_c w[c] * x[c], sum_c w[c]
        for (int c = 0; c < C; ++c) {
          const float x = col[c * 128 + tid];
          const float e = expf(x - m);
          const float wc = weights ? weights[c] : 1.f;
          sum += e;
          wx += wc * x;
          wsum += wc;
          if (BACKWARD) col[c * 128 + tid] = e;
        }
        const float wt = weights ? weights[t] : 1.f;
        const float smooth = eps / static_cast<float>(C);
        if (!BACKWARD) {
          const float logz = m + logf(sum);
          // (1-eps) * w[t] * (logZ - x[t]) + eps/C * sum_c w[c] * (logZ - x[c])
          const float xt = col[t * 128 + tid];
          my_loss = static_cast<double>((1.f - eps) * wt * (logz - xt)) +
                    static_cast<double>(smooth) * (static_cast<double>(wsum) * logz - static_cast<double>(wx));
          my_cnt = 1;
        } else {
          const float g = grad_out[0];
          const float a_sum = (1.f - eps) * wt + smooth * wsum;
          const float inv = 1.f / sum;
          for (int c = 0; c < C; ++c) {
            const float wc = weights ? weights[c] : 1.f;
            const float a = smooth * wc + (c == t ? (1.f - eps) * wt : 0.f);
            dlogits[(n * C + c) * HW + pix] = g * (a_sum * (col[c * 128 + tid] * inv) - a);
          }
        }
      }
    }
  }
  if (!BACKWARD) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
      my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, o);
    }
    if ((tid & 31) == 0) {
      s_loss[tid >> 5] = my_loss;
      s_cnt[tid >> 5] = my_cnt;
    }
    __syncthreads();
    if (tid == 0) {
      const int cnt = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
      if (cnt) {
        atomicAdd(loss_acc, s_loss[0] + s_loss[1] + s_loss[2] + s_loss[3]);
        atomicAdd(reinterpret_cast<unsigned long long*>(count_acc), static_cast<unsigned long long>(cnt & 0xFFFFF));
        if (cnt >> 20) atomicOr(reinterpret_cast<unsigned long long*>(count_acc), 1ull << 62);
      }
    }
  }
}

int check_common(const char* what, const float* logits, const void* target, int target_bytes, int N, int C, int H,
                 int W, float eps) {
  EB_REQUIRE(logits && target && N > 0 && C > 0 && H > 0 && W > 0, "%s: bad argument", what);
  EB_REQUIRE(target_bytes == 1 || target_bytes == 4 || target_bytes == 8, "%s: target elements of %d bytes", what,
             target_bytes);
  EB_REQUIRE(eps >= 0.f && eps <= 1.f, "%s: label smoothing %f outside [0,1]", what, eps);
  EB_REQUIRE(static_cast<size_t>(C) * 128 * 4 <= 200 * 1024, "%s: %d classes exceed the shared-memory column buffer", what,
             C);
  return 0;
}

}  // namespace

extern "C" int eb200_ce_loss_fwd(const float* logits, const void* target, int target_bytes, const float* weights,
                                 float label_smoothing, int N, int C, int H, int W, double* loss_acc,
                                 long long* count_acc, void* stream) {
  if (int rc = check_common("eb200_ce_loss_fwd", logits, target, target_bytes, N, C, H, W, label_smoothing)) return rc;
  EB_REQUIRE(loss_acc && count_acc, "eb200_ce_loss_fwd: missing accumulators");
  const long long HW = static_cast<long long>(H) * W, total = HW * N;
  const size_t smem = static_cast<size_t>(C) * 128 * 4;
  EB_CUDA(cudaMemsetAsync(loss_acc, 0, sizeof(double), STREAM));
  EB_CUDA(cudaMemsetAsync(count_acc, 0, sizeof(long long), STREAM));
  if (smem > 48 * 1024)
    EB_CUDA(cudaFuncSetAttribute(ce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ce_kernel<false><<<static_cast<int>((total + 127) / 128), 128, smem, STREAM>>>(
      logits, target, target_bytes, weights, label_smoothing, C, HW, total, nullptr, nullptr, loss_acc, count_acc);
  return eb::launch_check("ce_kernel<fwd>");
}

extern "C" int eb200_ce_loss_bwd(const float* logits, const void* target, int target_bytes, const float* weights,
                                 float label_smoothing, const float* grad_out, int N, int C, int H, int W,
                                 float* dlogits, void* stream) {
  if (int rc = check_common("eb200_ce_loss_bwd", logits, target, target_bytes, N, C, H, W, label_smoothing)) return rc;
  EB_REQUIRE(grad_out && dlogits, "eb200_ce_loss_bwd: missing grad_out / dlogits");
  const long long HW = static_cast<long long>(H) * W, total = HW * N;
  const size_t smem = static_cast<size_t>(C) * 128 * 4;
  if (smem > 48 * 1024)
    EB_CUDA(cudaFuncSetAttribute(ce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ce_kernel<true><<<static_cast<int>((total + 127) / 128), 128, smem, STREAM>>>(
      logits, target, target_bytes, weights, label_smoothing, C, HW, total, grad_out, dlogits, nullptr, nullptr);
  return eb::launch_check("ce_kernel<bwd>");
}

namespace {
// ------------------------------------------------------------------------------------------------------------------
// Masked regression losses of the instance / orientation task (MT/loss/mse.py:13-41, l1.py:13-41, vonmises.py:18-51 as
// MT/task_helper/instance.py:92-207 composes them):
//   kind 0 (MSE, centers)      loss = sum_p 1/C sum_c (x_pc * m_p - t_pc)^2        every pixel contributes
//   kind 1 (L1, offsets)       loss = sum_p 1/C sum_c |x_pc * m_p - t_pc|
//   kind 2 (von Mises, biternion orientations)  loss = sum_{p: m_p} 1 - exp(kappa * (sum_c x_pc t_pc - 1))
// and count = sum_p m_p (the "number of valid pixels" the reference fetches with .sum().cpu().item(): three host
// synchronisations per scale there, none here).  The reference multiplies the prediction by the mask, builds the
// elementwise loss tensor, reduces over channels and sums: 5-7 full-tensor passes plus their autograd backward; here
// one pass forward and one pass backward.  Element (n, c, p) of pred / target / dpred lives at n*sn + c*sc + p*sp.
template <int KIND, bool BACKWARD>
__global__ void __launch_bounds__(256) masked_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          const unsigned char* __restrict__ mask, int C, long long P,
                                                          long long total, long long sn, long long sc, long long sp,
                                                          float kappa, const float* __restrict__ grad_out,
                                                          float* __restrict__ dpred, double* __restrict__ loss_acc,
                                                          long long* __restrict__ count_acc) {
  __shared__ double s_loss[8];
  __shared__ int s_cnt[8];
  double my_loss = 0.0;
  int my_cnt = 0;
  const float g = BACKWARD ? grad_out[0] : 0.f;
  const float inv_c = 1.f / static_cast<float>(C);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    const bool m = mask ? mask[i] != 0 : true;
    const float mf = m ? 1.f : 0.f;
    const long long base = n * sn + p * sp;
    if (KIND == 2) {
      if (!m) {
        if (BACKWARD)
          for (int c = 0; c < C; ++c) dpred[base + c * sc] = 0.f;
        continue;
      }
      float dot = 0.f;
      for (int c = 0; c < C; ++c) dot = fmaf(pred[base + c * sc], target[base + c * sc], dot);
      const float e = expf(kappa * (dot - 1.f));
      if (!BACKWARD) {
        my_loss += static_cast<double>(1.f - e);
      } else {
        for (int c = 0; c < C; ++c) dpred[base + c * sc] = -g * kappa * e * target[base + c * sc];
      }
    } else {
      float acc = 0.f;
      for (int c = 0; c < C; ++c) {
        const float d = pred[base + c * sc] * mf - target[base + c * sc];
        if (!BACKWARD) {
          acc += KIND == 0 ? d * d : fabsf(d);
        } else {
          const float dl = KIND == 0 ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
          dpred[base + c * sc] = g * inv_c * dl * mf;
        }
      }
      if (!BACKWARD) my_loss += static_cast<double>(acc * inv_c);
    }
    my_cnt += m ? 1 : 0;
  }
  if (!BACKWARD) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
      my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_loss[threadIdx.x >> 5] = my_loss;
      s_cnt[threadIdx.x >> 5] = my_cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double l = 0.0;
      long long c = 0;
      for (int w = 0; w < 8; ++w) { l += s_loss[w]; c += s_cnt[w]; }
      atomicAdd(loss_acc, l);
      if (c) atomicAdd(reinterpret_cast<unsigned long long*>(count_acc), static_cast<unsigned long long>(c));
    }
  }
}

template <bool BACKWARD>
int launch_masked(int kind, const float* pred, const float* target, const unsigned char* mask, int N, int C, long long P,
                  long long sn, long long sc, long long sp, float kappa, const float* grad_out, float* dpred,
                  double* loss_acc, long long* count_acc, cudaStream_t st) {
  const long long total = static_cast<long long>(N) * P;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
#define EB_ML(K)                                                                                                       \
  masked_loss_kernel<K, BACKWARD><<<blocks, 256, 0, st>>>(pred, target, mask, C, P, total, sn, sc, sp, kappa, grad_out, \
                                                          dpred, loss_acc, count_acc)
  if (kind == 0) EB_ML(0);
  else if (kind == 1) EB_ML(1);
  else EB_ML(2);
#undef EB_ML
  return eb::launch_check("masked_loss_kernel");
}

}  // namespace

extern "C" int eb200_masked_loss_fwd(int kind, const float* pred, const float* target, const void* mask, int N, int C,
                                     long long P, long long sn, long long sc, long long sp, float kappa,
                                     double* loss_acc, long long* count_acc, void* stream) {
  EB_REQUIRE(kind >= 0 && kind <= 2 && pred && target && loss_acc && count_acc && N > 0 && C > 0 && C <= 8 && P > 0,
             "eb200_masked_loss_fwd: bad argument");
  EB_CUDA(cudaMemsetAsync(loss_acc, 0, sizeof(double), STREAM));
  EB_CUDA(cudaMemsetAsync(count_acc, 0, sizeof(long long), STREAM));
  return launch_masked<false>(kind, pred, target, static_cast<const unsigned char*>(mask), N, C, P, sn, sc, sp, kappa,
                              nullptr, nullptr, loss_acc, count_acc, STREAM);
}

extern "C" int eb200_masked_loss_bwd(int kind, const float* pred, const float* target, const void* mask, int N, int C,
                                     long long P, long long sn, long long sc, long long sp, float kappa,
                                     const float* grad_out, float* dpred, void* stream) {
  EB_REQUIRE(kind >= 0 && kind <= 2 && pred && target && grad_out && dpred && N > 0 && C > 0 && C <= 8 && P > 0,
             "eb200_masked_loss_bwd: bad argument");
  return launch_masked<true>(kind, pred, target, static_cast<const unsigned char*>(mask), N, C, P, sn, sc, sp, kappa,
                             grad_out, dpred, nullptr, nullptr, STREAM);
}
