// Normalisation and depthwise-convolution kernels on tokens-major / NHWC fp32 activations.
// All of these are HBM/L2-bound: one warp owns one token row (coalesced float4 along channels),
// reductions are warp shuffles, no shared-memory staging is needed because a row is read once.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "bwd.cuh"

namespace {

struct LnGroups {
  LnGroup g[TCX_MAX_GROUPS];
};

// One warp per row, row held in registers (NV float4 per lane), exact two-pass statistics.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnGroups gs, long long M, int C, float eps) {
  PDL_TOP();
  const LnGroup& g = gs.g[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = C >> 2;
  const float4* __restrict__ x = reinterpret_cast<const float4*>(g.x + row * C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int idx = lane + i * 32;
    v[i] = idx < nvec ? x[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  const float4* __restrict__ w = reinterpret_cast<const float4*>(g.w);
  const float4* __restrict__ b = reinterpret_cast<const float4*>(g.b);
  float4* __restrict__ y = reinterpret_cast<float4*>(g.y + row * C);
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      const float4 ww = w[idx], bb = b[idx];
      float4 o;
      o.x = (v[i].x - mean) * rstd * ww.x + bb.x;
      o.y = (v[i].y - mean) * rstd * ww.y + bb.y;
      o.z = (v[i].z - mean) * rstd * ww.z + bb.z;
      o.w = (v[i].w - mean) * rstd * ww.w + bb.w;
      y[idx] = o;
    }
  }
}

struct DwGroups {
  DwGroup g[TCX_MAX_GROUPS];
};

// Depthwise 3x3, pad 1, stride 1|2, NHWC. One thread = one output pixel x 4 channels.
template <int EPI>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwGroups gs, int B, int H, int W, int C, int stride,
                                                        int Ho, int Wo, BnParams bn) {
  PDL_TOP();
  const DwGroup& g = gs.g[blockIdx.y];
  const int cv = C >> 2;
  const long long total = (long long)B * Ho * Wo * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % cv);
  long long pix = idx / cv;
  const int wo = (int)(pix % Wo); pix /= Wo;
  const int ho = (int)(pix % Ho);
  const int b = (int)(pix / Ho);
  const int c = c4 * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g.b) acc = *reinterpret_cast<const float4*>(g.b + c);
  const float* __restrict__ wt = g.w + (long long)c * 9;
  float4 centre = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int ky = 0; ky < 3; ky++) {
    const int hi = ho * stride - 1 + ky;
    if (hi < 0 || hi >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; kx++) {
      const int wi = wo * stride - 1 + kx;
      if (wi < 0 || wi >= W) continue;
      const float4 xv = *reinterpret_cast<const float4*>(g.x + (((long long)b * H + hi) * W + wi) * C + c);
      const int t = ky * 3 + kx;
      acc.x = fmaf(xv.x, __ldg(wt + t), acc.x);
      acc.y = fmaf(xv.y, __ldg(wt + 9 + t), acc.y);
      acc.z = fmaf(xv.z, __ldg(wt + 18 + t), acc.z);
      acc.w = fmaf(xv.w, __ldg(wt + 27 + t), acc.w);
      if (EPI == DW_ADD_INPUT && ky == 1 && kx == 1) centre = xv;
    }
  }
  if (EPI == DW_ADD_INPUT) {
    acc.x += centre.x; acc.y += centre.y; acc.z += centre.z; acc.w += centre.w;
  }
  if (EPI == DW_BN_HS) {
    float s, t;
    bn_fold(bn, c + 0, s, t); acc.x = hardswish(acc.x * s + t);
    bn_fold(bn, c + 1, s, t); acc.y = hardswish(acc.y * s + t);
    bn_fold(bn, c + 2, s, t); acc.z = hardswish(acc.z * s + t);
    bn_fold(bn, c + 3, s, t); acc.w = hardswish(acc.w * s + t);
  }
  *reinterpret_cast<float4*>(g.y + (((long long)b * Ho + ho) * Wo + wo) * C + c) = acc;
}

struct MixMidGroups {
  MixMidGroup g[TCX_MAX_GROUPS];
};

// Mix-FFN middle: y = GELU(LN_C4(dw3x3(h) + b + h)) (reference MSTr.py:59). One warp per token; the
// C4-wide row stays in registers between the convolution and the LayerNorm.
template <int NV>
__global__ void __launch_bounds__(256) mixffn_mid_kernel(const MixMidGroups gs, int B, int H, int W, int C4, float eps) {
  PDL_TOP();
  const MixMidGroup& g = gs.g[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)B * H * W;
  if (tok >= total) return;
  const int wq = (int)(tok % W);
  const int hq = (int)((tok / W) % H);
  const int nvec = C4 >> 2;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int idx = lane + i * 32;
    v[i] = idx < nvec ? *reinterpret_cast<const float4*>(g.dwb + idx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int ky = 0; ky < 3; ky++) {
    const int hi = hq - 1 + ky;
    if (hi < 0 || hi >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; kx++) {
      const int wi = wq - 1 + kx;
      if (wi < 0 || wi >= W) continue;
      const float* __restrict__ src = g.h + (tok + (long long)(ky - 1) * W + (kx - 1)) * C4;
      const int t = ky * 3 + kx;
#pragma unroll
      for (int i = 0; i < NV; i++) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
          const float4 xv = *reinterpret_cast<const float4*>(src + idx * 4);
          const float* __restrict__ wt = g.dww + (long long)idx * 36 + t;
          float w0 = __ldg(wt), w1 = __ldg(wt + 9), w2 = __ldg(wt + 18), w3 = __ldg(wt + 27);
          if (t == 4) { w0 += 1.f; w1 += 1.f; w2 += 1.f; w3 += 1.f; }  // + h (skip)
          v[i].x = fmaf(xv.x, w0, v[i].x);
          v[i].y = fmaf(xv.y, w1, v[i].y);
          v[i].z = fmaf(xv.z, w2, v[i].z);
          v[i].w = fmaf(xv.w, w3, v[i].w);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);  // padded lanes hold 0
  const float mean = warp_sum(s) / (float)C4;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (lane + i * 32 < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C4 + eps);
  float* __restrict__ y = g.y + tok * C4;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      const float4 ww = *reinterpret_cast<const float4*>(g.lnw + idx * 4);
      const float4 bb = *reinterpret_cast<const float4*>(g.lnb + idx * 4);
      float4 o;
      o.x = gelu_erf((v[i].x - mean) * rstd * ww.x + bb.x);
      o.y = gelu_erf((v[i].y - mean) * rstd * ww.y + bb.y);
      o.z = gelu_erf((v[i].z - mean) * rstd * ww.z + bb.z);
      o.w = gelu_erf((v[i].w - mean) * rstd * ww.w + bb.w);
      *reinterpret_cast<float4*>(y + idx * 4) = o;
    }
  }
}

// fp32 -> fp16 (round to nearest, saturating), 8 elements per thread
__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  PDL_TOP();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
    __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y),
                    __floats2half2_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<uint4*>(h);
  } else {
    for (long long j = i; j < n; j++) dst[j] = __float2half_rn(src[j]);
  }
}

// fp32 -> fp16(x * scale), clamped to the finite fp16 range (a gradient operand never becomes inf)
__global__ void __launch_bounds__(256) f32_to_f16_scaled_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n,
                                                               float scale) {
  PDL_TOP();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  auto cv = [scale](float v) { return fminf(fmaxf(v * scale, -65504.f), 65504.f); };
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
    __half2 h[4] = {__floats2half2_rn(cv(a.x), cv(a.y)), __floats2half2_rn(cv(a.z), cv(a.w)), __floats2half2_rn(cv(b.x), cv(b.y)),
                    __floats2half2_rn(cv(b.z), cv(b.w))};
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<uint4*>(h);
  } else {
    for (long long j = i; j < n; j++) dst[j] = __float2half_rn(cv(src[j]));
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  PDL_TOP();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
    __nv_bfloat162 h[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w), __floats2bfloat162_rn(b.x, b.y),
                           __floats2bfloat162_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<uint4*>(h);
  } else {
    for (long long j = i; j < n; j++) dst[j] = __float2bfloat16_rn(src[j]);
  }
}

__global__ void __launch_bounds__(256) f16_to_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, long long n) {
  PDL_TOP();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 raw = *reinterpret_cast<const uint4*>(src + i);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
    const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
    *reinterpret_cast<float4*>(dst + i) = make_float4(a.x, a.y, b.x, b.y);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(c.x, c.y, d.x, d.y);
  } else {
    for (long long j = i; j < n; j++) dst[j] = __half2float(src[j]);
  }
}

// up to three independent tensors -> bf16 in one launch (the operands of one Linear backward: dy fp32, x fp16 | fp32, w fp32)
__global__ void __launch_bounds__(256) to_bf16_multi_kernel(const CvtSegs segs) {
  PDL_TOP();
  long long chunk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int k = 0;
  while (k < segs.n && chunk >= segs.chunks[k]) { chunk -= segs.chunks[k]; k++; }
  if (k >= segs.n) return;
  const long long i = chunk * 8, n = segs.count[k];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(segs.dst[k]);
  if (segs.src_f16[k]) {
    const __half* src = reinterpret_cast<const __half*>(segs.src[k]);
    if (i + 8 <= n) {
      const uint4 raw = *reinterpret_cast<const uint4*>(src + i);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
      __nv_bfloat162 o[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { const float2 f = __half22float2(h[q]); o[q] = __floats2bfloat162_rn(f.x, f.y); }
      *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<uint4*>(o);
    } else {
      for (long long j = i; j < n; j++) dst[j] = __float2bfloat16_rn(__half2float(src[j]));
    }
  } else {
    const float* src = reinterpret_cast<const float*>(segs.src[k]);
    if (i + 8 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
      __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w), __floats2bfloat162_rn(b.x, b.y),
                             __floats2bfloat162_rn(b.z, b.w)};
      *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<uint4*>(o);
    } else {
      for (long long j = i; j < n; j++) dst[j] = __float2bfloat16_rn(src[j]);
    }
  }
}

}  // namespace

int launch_f16_to_f32(const __half* src, float* dst, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  TCX_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "f16_to_f32: pointers must be 16-byte aligned");
  const long long threads = (n + 7) / 8;
  tcx_launch_chain(f16_to_f32_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, src, dst, n);
  return tcx_check_launch("f16_to_f32");
}

int launch_to_bf16_multi(CvtSegs segs, cudaStream_t st) {
  long long total = 0;
  for (int k = 0; k < segs.n; k++) {
    TCX_REQUIRE((((uintptr_t)segs.src[k] | (uintptr_t)segs.dst[k]) & 15) == 0, "to_bf16: pointers must be 16-byte aligned");
    segs.chunks[k] = (segs.count[k] + 7) / 8;
    total += segs.chunks[k];
  }
  if (total == 0) return 0;
  tcx_launch_chain(to_bf16_multi_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, segs);
  return tcx_check_launch("to_bf16_multi");
}

int launch_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  TCX_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "f32_to_bf16: pointers must be 16-byte aligned");
  const long long threads = (n + 7) / 8;
  tcx_launch_chain(f32_to_bf16_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  return tcx_check_launch("f32_to_bf16");
}

int launch_f32_to_f16_scaled(const float* src, __half* dst, long long n, float scale, cudaStream_t st) {
  if (n <= 0) return 0;
  TCX_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "f32_to_f16_scaled: pointers must be 16-byte aligned");
  const long long threads = (n + 7) / 8;
  tcx_launch_chain(f32_to_f16_scaled_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, src, dst, n, scale);
  return tcx_check_launch("f32_to_f16_scaled");
}

int launch_f32_to_f16(const float* src, void* dst, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  TCX_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "f32_to_f16: pointers must be 16-byte aligned");
  const long long threads = (n + 7) / 8;
  tcx_launch_chain(f32_to_f16_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, src, reinterpret_cast<__half*>(dst), n);
  return tcx_check_launch("f32_to_f16");
}

#define DISPATCH_NV(C, CALL)                                         \
  do {                                                               \
    const int _nv = ((C) / 4 + 31) / 32;                             \
    if (_nv <= 1) { CALL(1); }                                       \
    else if (_nv <= 2) { CALL(2); }                                  \
    else if (_nv <= 3) { CALL(3); }                                  \
    else if (_nv <= 4) { CALL(4); }                                  \
    else if (_nv <= 8) { CALL(8); }                                  \
    else if (_nv <= 10) { CALL(10); }                                \
    else if (_nv <= 16) { CALL(16); }                                \
    else { tcx_set_error("row width %d > 2048 not supported", (C)); return -1; } \
  } while (0)

int launch_layernorm_grouped(const LnGroup* g, int groups, long long M, int C, float eps, cudaStream_t st) {
  TCX_REQUIRE(C % 4 == 0 && groups >= 1 && groups <= TCX_MAX_GROUPS, "layernorm: C %% 4 != 0 or bad groups");
  if (M == 0) return 0;
  LnGroups gs{};
  for (int i = 0; i < groups; i++) gs.g[i] = g[i];
  dim3 grid((unsigned)((M + 7) / 8), groups);
#define CALL(NV) tcx_launch_chain(layernorm_kernel<NV>, dim3(grid), dim3(256), 0, st, gs, M, C, eps)
  DISPATCH_NV(C, CALL);
#undef CALL
  return tcx_check_launch("layernorm");
}

int launch_layernorm(const float* x, const float* w, const float* b, float* y, long long M, int C, float eps,
                     cudaStream_t st) {
  LnGroup g{x, w, b, y};
  return launch_layernorm_grouped(&g, 1, M, C, eps, st);
}

int launch_dwconv3x3(const DwGroup* g, int groups, int B, int H, int W, int C, int stride, int epi, BnParams bn,
                     cudaStream_t st) {
  TCX_REQUIRE(C % 4 == 0 && (stride == 1 || stride == 2), "dwconv3x3: C %% 4 != 0 or bad stride");
  DwGroups gs{};
  for (int i = 0; i < groups; i++) gs.g[i] = g[i];
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = (long long)B * Ho * Wo * (C / 4);
  dim3 grid((unsigned)((total + 255) / 256), groups);
  if (epi == DW_PLAIN) tcx_launch_chain(dwconv3x3_kernel<DW_PLAIN>, dim3(grid), dim3(256), 0, st, gs, B, H, W, C, stride, Ho, Wo, bn);
  else if (epi == DW_ADD_INPUT) tcx_launch_chain(dwconv3x3_kernel<DW_ADD_INPUT>, dim3(grid), dim3(256), 0, st, gs, B, H, W, C, stride, Ho, Wo, bn);
  else tcx_launch_chain(dwconv3x3_kernel<DW_BN_HS>, dim3(grid), dim3(256), 0, st, gs, B, H, W, C, stride, Ho, Wo, bn);
  return tcx_check_launch("dwconv3x3");
}

int launch_mixffn_mid(const MixMidGroup* g, int groups, int B, int H, int W, int C4, float eps, cudaStream_t st) {
  TCX_REQUIRE(C4 % 4 == 0, "mixffn_mid: C4 %% 4 != 0");
  MixMidGroups gs{};
  for (int i = 0; i < groups; i++) gs.g[i] = g[i];
  const long long total = (long long)B * H * W;
  dim3 grid((unsigned)((total + 7) / 8), groups);
  ProfScope prof("mixffn_mid", st);
#define CALL(NV) tcx_launch_chain(mixffn_mid_kernel<NV>, dim3(grid), dim3(256), 0, st, gs, B, H, W, C4, eps)
  DISPATCH_NV(C4, CALL);
#undef CALL
  return tcx_check_launch("mixffn_mid");
}
