#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

#ifdef __CUDACC__
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the fp16
// output quantum).  16 instructions, no branches: MUFU.RCP and MUFU.EX2 are issued directly (`__frcp_rn` / `__expf` expand
// to range checks, a Newton step and a divergent slow-path call: 31 instructions per element in the round-1c SASS).
__device__ __forceinline__ float tcx_gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(x * (x * -0.72134752044448170f)));     // exp(-z^2) = 2^(-x^2 log2(e) / 2)
  const float e = p * t * ex;                        // 1 - erf(z), z >= 0
  const float h = 0.5f * x;
  return x >= 0.f ? fmaf(-h, e, x) : h * e;          // x>=0: 0.5x(2-e) ; x<0: 0.5x(1-(1-e)) = 0.5 x e
}
#endif

struct DwLnGroup {
  const void* x;      // [B,H,W,C] fp32 or fp16
  const float* dww;   // [C,1,3,3]
  const float* dwb;   // [C] or null
  const float* lnw;   // [C]
  const float* lnb;   // [C]
  float* u;           // optional fp32 output: dw3x3(x) + b + x
  __half* y;          // fp16 output: [GELU](LayerNorm(u))
};
struct DwLnArgs {
  DwLnGroup g[TCX_MAX_GROUPS];
  int B, H, W, C;
  float eps;
  int tpw;            // tokens per warp slot (set by the launcher)
  int gelu;
};
int launch_dwln(DwLnArgs a, int groups, bool in16, cudaStream_t st);

struct Ln16Group {
  const float* x;     // [M,C] fp32
  const float* w;
  const float* b;
  __half* y16;        // optional
  float* y32;         // optional
};
struct Ln16Args {
  Ln16Group g[TCX_MAX_GROUPS];
  long long M;
  int C;
  float eps;
  int tpw;
};
int launch_ln16(Ln16Args a, int groups, cudaStream_t st);

struct Mb16Args {
  int B, H, W, C, heads;
  float scale;
  const __half* qkv[TCX_MAX_GROUPS];    // [B*N][3C]
  float* ctx[TCX_MAX_GROUPS];           // [B][heads][Ch][Ch]
  __half* out[TCX_MAX_GROUPS];          // [B*N][C]
  const float* cw[TCX_MAX_GROUPS][3];   // crpe filters 3x3 / 5x5 / 7x7
  const float* cb[TCX_MAX_GROUPS][3];
};
int launch_mb_attention16(const Mb16Args& a, int groups, cudaStream_t st);

// fused Mix-FFN tail: y = res + fc2(GELU(LN(dw3x3(h) + b + h))) for C4 in {256, 512} (mixtail.cu)
struct MixTailDesc {
  const void* h;      // [B*N][C4] fp16 fc1 output
  const float* dww;   // [C4,1,3,3]
  const float* dwb;   // [C4] or null
  const float* lnw;   // [C4]
  const float* lnb;   // [C4]
  const void* w2;     // prepared fp16 fc2 weight [C][C4]
  const float* b2;    // [C]
  const float* res;   // fp32 residual rows or null
  float* y;           // fp32 output rows
};
bool mixtail_eligible(int groups, int C4, long long tokens);
int launch_mixtail(const MixTailDesc* d, int groups, int B, int H, int W, int C4, float eps, long long res_bs, long long y_bs,
                   cudaStream_t st);
