// Fused multi-tensor optimizer step: emsanet/optimizer.py:29-59 builds torch.optim.SGD
// (momentum, nesterov=True, weight decay) / Adam / AdamW over all 675 parameter tensors and main.py:597-599 calls
// optimizer.step() after loss.backward(); the engine then re-lays-out every conv weight for the tensor cores.
// Here ONE launch does both: every block updates a slice of one parameter from the flat fp32 gradient buffer (fp32
// master parameter and optimizer state in place) and, for tensor-core conv weights, writes the two bf16 operand layouts
// ([tap][co][ci] for the forward pass, [tap][ci][co] for the data gradient) from the freshly updated values — the
// separate eb200_pack_conv_weights_batched launch and its re-read of all weights disappear.
//
// Arithmetic follows torch.optim's multi-tensor (foreach) implementations operation by operation, including where
// torch rounds twice (buf.mul_(momentum).add_(grad)) and where its `a + alpha * b` kernels contract to one fma, so
// that fp32 master parameters and optimizer state stay BIT-identical to torch.optim.SGD / Adam / AdamW over many steps
// (tests/test_optim.py, scripts/probe_sgd_bits.py, scripts/probe_adam.py); the Adam scalars (1 - beta, lr /
// bias_correction1, 1 - lr * weight_decay) arrive pre-computed in double precision, as torch.optim computes them.
// Compiled without --use_fast_math.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/emsanet_b200.h"
#include "common.h"

#define STREAM static_cast<cudaStream_t>(stream)

namespace {

struct Upd {
  int kind;
  float lr, mom, beta2, eps, wd, bc1, bc2_sqrt, om1, om2, step_size, decay;
  int nesterov, first, flags;
};

__device__ __forceinline__ float axpy(float a, float alpha, float b, bool fma) {   // a + alpha * b
  return fma ? __fmaf_rn(alpha, b, a) : __fadd_rn(a, __fmul_rn(alpha, b));
}

// one element: returns the new parameter value; m / v updated in place
__device__ __forceinline__ float update(const Upd& u, float p, float g, float* m, float* v) {
  if (u.kind == EB200_OPT_SGD) {
    if (u.wd != 0.f) g = axpy(g, u.wd, p, !(u.flags & 1));                 // grad.add(param, alpha=weight_decay)
    if (u.mom != 0.f) {
      float buf;
      if (u.first) buf = g;                                                  // momentum_buffer = clone(grad)
      else buf = __fadd_rn(__fmul_rn(*m, u.mom), g);                         // buf.mul_(momentum).add_(grad)
      *m = buf;
      g = u.nesterov ? axpy(g, u.mom, buf, !(u.flags & 2)) : buf;            // grad.add(buf, alpha=momentum)
    }
    return axpy(p, -u.lr, g, !(u.flags & 4));                                // param.add_(grad, alpha=-lr)
  }
  if (u.kind == EB200_OPT_ADAMW) p = __fmul_rn(p, u.decay);                  // param.mul_(1 - lr * weight_decay)
  else if (u.wd != 0.f) g = __fmaf_rn(u.wd, p, g);                           // Adam: grad.add(param, alpha=wd)
  float mm = u.first ? 0.f : *m, vv = u.first ? 0.f : *v;
  mm = __fmaf_rn(u.om1, __fsub_rn(g, mm), mm);                               // exp_avg.lerp_(grad, 1 - beta1)
  vv = __fmaf_rn(u.om2, __fmul_rn(g, g), __fmul_rn(vv, u.beta2));            // mul_(beta2).addcmul_(g, g, 1 - beta2)
  *m = mm;
  *v = vv;
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), u.bc2_sqrt), u.eps);
  return __fmaf_rn(-u.step_size, __fdiv_rn(mm, denom), p);                   // addcdiv_(exp_avg, denom, -step_size)
}

__device__ __forceinline__ Upd load_hyper(const eb200_optim_hyper* h) {
  Upd u;
  u.kind = h->kind; u.lr = h->lr; u.mom = h->momentum; u.beta2 = h->beta2; u.eps = h->eps; u.wd = h->weight_decay;
  u.bc1 = h->bias_correction1; u.bc2_sqrt = h->bias_correction2_sqrt; u.nesterov = h->nesterov;
  u.first = h->step <= 1; u.flags = h->flags;
  u.om1 = h->one_minus_beta1; u.om2 = h->one_minus_beta2; u.step_size = h->step_size; u.decay = h->decay;
  return u;
}

constexpr int kChunk = 4096;   // elements of a plain parameter per block

// Block b works on entry block_entry[b]:
//   plain parameter        : elements [kChunk * block_start[b], +kChunk)
//   tensor-core conv weight: the 32 x 32 (co, ci) tile (block_start[b] & 0xffff, block_start[b] >> 16) * 32, all taps —
//                            updated in registers / shared memory, then written to both bf16 layouts
__global__ void __launch_bounds__(256) optim_step_kernel(const eb200_optim_entry* __restrict__ entries,
                                                         const eb200_pack_entry* __restrict__ packs,
                                                         const int* __restrict__ block_entry,
                                                         const int* __restrict__ block_start,
                                                         const eb200_optim_hyper hyper) {
  __shared__ float tile[32][32 * 9 + 1];
  const Upd u = load_hyper(&hyper);
  const eb200_optim_entry e = entries[block_entry[blockIdx.x]];
  if (e.pack < 0) {
    const long long i0 = static_cast<long long>(block_start[blockIdx.x]) * kChunk;
    const long long i1 = i0 + kChunk < e.numel ? i0 + kChunk : e.numel;
    for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
      float m = 0.f, v = 0.f;
      if (e.m && !u.first) m = e.m[i];
      if (e.v && !u.first) v = e.v[i];
      const float p = update(u, e.p[i], e.g[i], &m, &v);
      e.p[i] = p;
      if (e.m) e.m[i] = m;
      if (e.v) e.v[i] = v;
    }
    return;
  }
  const eb200_pack_entry pk = packs[e.pack];
  const int co0 = (block_start[blockIdx.x] & 0xffff) * 32, ci0 = (block_start[blockIdx.x] >> 16) * 32;
  const int nco = min(32, pk.cout - co0), nci = min(32, pk.cin - ci0);
  const int rowlen = nci * pk.taps;                       // contiguous floats of one co row of the tile
  for (int i = threadIdx.x; i < nco * rowlen; i += 256) {
    const int r = i / rowlen, c = i - r * rowlen;
    const size_t off = (static_cast<size_t>(co0 + r) * pk.cin + ci0) * pk.taps + c;
    float m = 0.f, v = 0.f;
    if (e.m && !u.first) m = e.m[off];
    if (e.v && !u.first) v = e.v[off];
    const float p = update(u, e.p[off], e.g[off], &m, &v);
    e.p[off] = p;
    if (e.m) e.m[off] = m;
    if (e.v) e.v[off] = v;
    tile[r][c] = p;
  }
  __syncthreads();
  __nv_bfloat16* fwd = static_cast<__nv_bfloat16*>(pk.fwd);
  __nv_bfloat16* bwd = static_cast<__nv_bfloat16*>(pk.bwd);
  const int total = pk.taps * 32 * 32;
  for (int i = threadIdx.x; i < total; i += 256) {       // forward layout: ci fastest
    const int ci = i & 31, co = (i >> 5) & 31, t = i >> 10;
    if (co < nco && ci < nci)
      fwd[(static_cast<size_t>(t) * pk.fwd_rows + co0 + co + pk.co_off) * pk.fwd_cols + ci0 + ci + pk.ci_off] =
          __float2bfloat16(tile[co][ci * pk.taps + t]);
  }
  if (bwd) {
    for (int i = threadIdx.x; i < total; i += 256) {     // transposed layout: co fastest
      const int co = i & 31, ci = (i >> 5) & 31, t = i >> 10;
      if (co < nco && ci < nci)
        bwd[(static_cast<size_t>(t) * pk.bwd_rows + ci0 + ci + pk.ci_off) * pk.bwd_cols + co0 + co + pk.co_off] =
            __float2bfloat16(tile[co][ci * pk.taps + t]);
    }
  }
}

}  // namespace

extern "C" int eb200_optim_chunk(void) { return kChunk; }

extern "C" int eb200_optim_step(const eb200_optim_entry* entries_dev, const eb200_pack_entry* packs_dev,
                                const int* block_entry_dev, const int* block_start_dev, int nblocks,
                                const eb200_optim_hyper* hyper, void* stream) {
  EB_REQUIRE(entries_dev && block_entry_dev && block_start_dev && hyper && nblocks > 0,
             "eb200_optim_step: bad argument");
  optim_step_kernel<<<nblocks, 256, 0, STREAM>>>(entries_dev, packs_dev, block_entry_dev, block_start_dev, *hyper);
  return eb::launch_check("optim_step_kernel");
}
