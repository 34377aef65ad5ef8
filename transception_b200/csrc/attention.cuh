#pragma once
#include "common.cuh"

// element (b, c, n) of K/Q/V lives at ptr + b*sb + c*sc + n*sn
struct EaView {
  const float* k;
  const float* q;
  const float* v;
  long long sb, sc, sn;
};
size_t ea_workspace_floats(int B, int N, int C);
int launch_ea_context(const EaView& v, bool reinterpret, int B, int N, int C, float* ws, float* ctxT, cudaStream_t st);
int launch_ea_qsoftmax(const EaView& v, bool reinterpret, int B, int N, int C, float* dst, cudaStream_t st);

struct MbAttnArgs {
  int groups, B, H, W, C, heads;
  float scale;
  const float* qkv[TCX_MAX_GROUPS];
  float* ctx[TCX_MAX_GROUPS];
  float* out[TCX_MAX_GROUPS];
  const float* cw[TCX_MAX_GROUPS][3];
  const float* cb[TCX_MAX_GROUPS][3];
};
int launch_mb_attention(const MbAttnArgs& a, cudaStream_t st);

int launch_flash_ffma(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, cudaStream_t st);
size_t flash_tc_workspace_bytes(int B, int Nk);
int launch_flash_tc(const float* q, const float* kv, float* out, int B, int Nq, int Nk, float scale, void* ws,
                    cudaStream_t st, float* lse = nullptr);
// flash backward (flash_bwd.cu): lse = the row log2-sum-exp written by launch_flash_tc
size_t flash_bwd_workspace_floats(int B, int Nq, int Nk);
int launch_flash_bwd(const float* q, const float* kv, const float* out, const float* lse, const float* dout, float scale, float* dq,
                     float* dkv, int B, int Nq, int Nk, float* ws, cudaStream_t st);
int launch_flash_tc16(const void* q16, const void* kv16, void* out16, int B, int Nq, int Nk, float scale, void* ws,
                      cudaStream_t st);
bool flash_tc_enabled();
