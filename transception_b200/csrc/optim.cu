// Multi-tensor optimizer step of the training row (trainer.py:125,148: optim.SGD(momentum, weight_decay) after an optional
// clip_grad_norm_), as three launches over a device-resident tensor table instead of ~100 ATen multi_tensor_apply launches
// and 238 per-weight fp32 -> fp16 conversions per step:
//   1. mt_gather      the used gradients -> one flat fp32 bucket (multi-GPU only: the all-reduce operand, no torch.cat)
//   2. mt_sqnorm      sum of squares of all gradients, ordered (block partials -> the last block folds them in index order)
//                     -> clip coefficient min(1, max_norm / (norm + 1e-6)) as a device scalar
//   3. mt_sgd         g = coef * grad + wd * p;  buf = momentum * buf + g;  p -= lr * buf;  w16 = half(p) where the forward keeps a
//                     prepared fp16 copy.  lr and coef are read from device memory: the per-iteration learning-rate schedule of
//                     trainer.py:151-153 is a 4-byte host-to-device copy, not a graph re-capture.
// A block handles one chunk of MT_CHUNK elements of one tensor (tables built once by the caller: tensor pointers, element counts,
// and the (tensor, chunk) pair of every block).  No atomics on data: bit-reproducible.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int MT_CHUNK = 4096;
constexpr int MT_THREADS = 256;

__device__ unsigned g_mt_ticket = 0;

__global__ void __launch_bounds__(MT_THREADS) mt_gather_kernel(const float* const* __restrict__ src, const long long* __restrict__ numel,
                                                               const long long* __restrict__ offset, const int2* __restrict__ blocks,
                                                               float* __restrict__ flat) {
  PDL_TOP();
  const int2 bk = blocks[blockIdx.x];
  const long long n = numel[bk.x], i0 = (long long)bk.y * MT_CHUNK;
  const float* s = src[bk.x];
  float* d = flat + offset[bk.x];
  for (long long i = i0 + threadIdx.x; i < n && i < i0 + MT_CHUNK; i += MT_THREADS) d[i] = s[i];
}

__global__ void __launch_bounds__(MT_THREADS) mt_sqnorm_kernel(const float* const* __restrict__ g, const long long* __restrict__ numel,
                                                               const int2* __restrict__ blocks, int nblocks, float* __restrict__ part,
                                                               float max_norm, float* __restrict__ out /* [norm, coef] */) {
  PDL_TOP();
  __shared__ float sm[MT_THREADS];
  __shared__ bool last;
  const int2 bk = blocks[blockIdx.x];
  const long long n = numel[bk.x], i0 = (long long)bk.y * MT_CHUNK;
  const float* s = g[bk.x];
  float acc = 0.f;
  for (long long i = i0 + threadIdx.x; i < n && i < i0 + MT_CHUNK; i += MT_THREADS) acc = fmaf(s[i], s[i], acc);
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = MT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[blockIdx.x] = sm[0];
    __threadfence();
    const unsigned t = atomicAdd(&g_mt_ticket, 1u);
    last = t == (unsigned)(nblocks - 1);
    if (last) g_mt_ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block folds the partials in index order: thread t sums partials t, t + 256, ... then a fixed tree
  double a = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += MT_THREADS) a += (double)__ldcg(part + i);
  __shared__ double sd[MT_THREADS];
  sd[threadIdx.x] = a;
  __syncthreads();
  for (int o = MT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sd[threadIdx.x] += sd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(sd[0]);
    out[0] = norm;
    out[1] = max_norm > 0.f ? fminf(1.0f, max_norm / (norm + 1e-6f)) : 1.0f;
  }
}

__global__ void __launch_bounds__(MT_THREADS) mt_sgd_kernel(float* const* __restrict__ p, const float* const* __restrict__ g,
                                                            float* const* __restrict__ buf, __half* const* __restrict__ w16,
                                                            const long long* __restrict__ numel, const int2* __restrict__ blocks,
                                                            const float* __restrict__ lr_ptr, const float* __restrict__ coef_ptr,
                                                            float momentum, float wd) {
  PDL_TOP();
  const int2 bk = blocks[blockIdx.x];
  const long long n = numel[bk.x], i0 = (long long)bk.y * MT_CHUNK;
  float* pp = p[bk.x];
  const float* gg = g[bk.x];
  float* bb = buf[bk.x];
  __half* hh = w16[bk.x];
  const float lr = *lr_ptr, coef = coef_ptr ? *coef_ptr : 1.0f;
  for (long long i = i0 + threadIdx.x; i < n && i < i0 + MT_CHUNK; i += MT_THREADS) {
    const float w = pp[i];
    const float d = fmaf(wd, w, gg[i] * coef);
    const float b = fmaf(momentum, bb[i], d);
    const float nw = fmaf(-lr, b, w);
    bb[i] = b;
    pp[i] = nw;
    if (hh) hh[i] = __float2half_rn(nw);
  }
}

}  // namespace

extern "C" {

int tcx_mt_chunk(void) { return MT_CHUNK; }

int tcx_mt_gather(const void* src_ptrs, const void* numel, const void* offsets, const void* blocks, int nblocks, float* flat, void* stream) {
  TCX_REQUIRE(src_ptrs && numel && offsets && blocks && flat, "mt_gather: null pointer");
  if (nblocks <= 0) return 0;
  tcx_launch_chain(mt_gather_kernel, dim3(nblocks), dim3(MT_THREADS), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const float* const*>(src_ptrs), reinterpret_cast<const long long*>(numel), reinterpret_cast<const long long*>(offsets),
      reinterpret_cast<const int2*>(blocks), flat);
  return tcx_check_launch("mt_gather");
}

int tcx_mt_sqnorm(const void* grad_ptrs, const void* numel, const void* blocks, int nblocks, float* part, float max_norm, float* out,
                  void* stream) {
  TCX_REQUIRE(grad_ptrs && numel && blocks && part && out, "mt_sqnorm: null pointer");
  if (nblocks <= 0) return 0;
  tcx_launch_chain(mt_sqnorm_kernel, dim3(nblocks), dim3(MT_THREADS), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const float* const*>(grad_ptrs), reinterpret_cast<const long long*>(numel), reinterpret_cast<const int2*>(blocks), nblocks,
      part, max_norm, out);
  return tcx_check_launch("mt_sqnorm");
}

int tcx_mt_sgd(const void* param_ptrs, const void* grad_ptrs, const void* buf_ptrs, const void* w16_ptrs, const void* numel,
               const void* blocks, int nblocks, const float* lr, const float* coef, float momentum, float weight_decay, void* stream) {
  TCX_REQUIRE(param_ptrs && grad_ptrs && buf_ptrs && w16_ptrs && numel && blocks && lr, "mt_sgd: null pointer");
  if (nblocks <= 0) return 0;
  tcx_launch_chain(mt_sgd_kernel, dim3(nblocks), dim3(MT_THREADS), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<float* const*>(const_cast<void*>(param_ptrs)), reinterpret_cast<const float* const*>(grad_ptrs),
      reinterpret_cast<float* const*>(const_cast<void*>(buf_ptrs)), reinterpret_cast<__half* const*>(const_cast<void*>(w16_ptrs)),
      reinterpret_cast<const long long*>(numel), reinterpret_cast<const int2*>(blocks), lr, coef, momentum, weight_decay);
  return tcx_check_launch("mt_sgd");
}

}  // extern "C"
