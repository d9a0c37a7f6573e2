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
        atomicAdd(reinterpret_cast<unsigned long long*>(count_acc), static_cast<unsigned long long>(cnt));
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
